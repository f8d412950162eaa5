"""Phase timing of the tile kernel (needs a -DB200COMP_PROFILE=1 build selected with B200COMP_LIB)."""
import ctypes, sys
sys.path.insert(0, '.')
import numpy as np, torch
from image_transformation_b200 import _native, synth
from image_transformation_b200.batch import CutoutPool, CompositeBatch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
pool_np = synth.workload_pool("c3_4k_20obj")
sizes = {k: (v.shape[1], v.shape[0]) for k, v in pool_np.items()}
pool = CutoutPool(pool_np)
pls = [synth.workload_placements("c3_4k_20obj", sizes, i) for i in range(n)]
bg = torch.full((n, 2160, 3840, 4), 200, dtype=torch.uint8, device="cuda")
b = CompositeBatch(pool, [(3840, 2160)] * n, pls, backgrounds=bg)
for _ in range(3): b.run()
b.check()
lib = _native.lib()
out = (ctypes.c_ulonglong * 16)()
lib.b200comp_debug_profile_(out)
b.run(); b.check()
assert lib.b200comp_debug_profile_(out) == 0, "not a profiling build"
names = ["command block wait", "tile wait (background)", "tile fill (solid / plain loads)", "step decode + coefficient rows", "patch chunk wait", "H pass", "V pass + over", "identity / chunks passed on", "tile end", "loop overhead"]
tot = sum(out[i] for i in range(10))
for i in range(10):
    print(f"{names[i]:28s} {out[i] / tot * 100:6.2f} %   {out[i] / 296 / 8 / 1e3:9.1f} kcycles per warp")
print("total kcycles per compute warp", tot / 296 / 8 / 1e3)
