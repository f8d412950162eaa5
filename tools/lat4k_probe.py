import os, sys, time
sys.path.insert(0, '.')
import numpy as np
from PIL import Image
from image_transformation_b200 import compositor, synth, _native
pool = synth.workload_pool("c3_4k_20obj")
sizes = {k: (v.shape[1], v.shape[0]) for k, v in pool.items()}
pls = synth.workload_placements("c3_4k_20obj", sizes, 0)
bg = Image.new("RGBA", (3840, 2160), (38, 73, 115, 255))
objs = {k: Image.fromarray(v, "RGBA").copy() for k, v in pool.items()}
for _ in range(3): compositor.composite(bg, objs, pls)
t = time.perf_counter(); a = _native.rgba_array(bg); print("bg rgba_array ms", (time.perf_counter() - t) * 1e3, file=sys.stderr)
used = sorted({int(p["object_id"]) for p in pls})
t = time.perf_counter(); arrs = [_native.rgba_array(objs[k]) for k in used]; print("cutouts rgba_array ms", (time.perf_counter() - t) * 1e3, sum(x.nbytes for x in arrs) / 1e6, "MB", file=sys.stderr)
os.environ["B200COMP_TRACE"] = "1"
t = time.perf_counter(); compositor.composite(bg, objs, pls); print("python total ms", (time.perf_counter() - t) * 1e3, file=sys.stderr)
