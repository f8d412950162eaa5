import sys, time, os
sys.path.insert(0,'.')
import torch
from image_transformation_b200 import batch as B, synth
import bench
n=int(sys.argv[1])
pool, canvases, placements = bench.build_inputs("c3_4k_20obj", 0, n)
dpool = B.CutoutPool(pool, torch.device("cuda"))
out = torch.empty(n*3840*2160*4, dtype=torch.uint8, device="cuda")
t0=time.perf_counter()
cb = B.CompositeBatch(dpool, canvases, placements, solid=(1,2,3,255), out=out, host_threads=16)
torch.cuda.synchronize()
print("python+C create", time.perf_counter()-t0, "C only", cb.plan_create_s)
cb.run(); cb.check()
os.environ["B200COMP_TRACE_PLAN"]="1"
for i in range(3):
    cb.recreate(); print("recreate C", cb.plan_create_s)
