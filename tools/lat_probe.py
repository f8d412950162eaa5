import os, sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from PIL import Image
import golden_io as G
from image_transformation_b200 import compositor
bg, objs, pls, exp = G.case("c1_squarespace_1x1")
bg_i = Image.fromarray(bg, "RGBA"); objs_i = {k: Image.fromarray(v, "RGBA") for k, v in objs.items()}
for _ in range(5): compositor.composite(bg_i, objs_i, pls)
os.environ["B200COMP_TRACE"] = "1"
t = time.perf_counter(); compositor.composite(bg_i, objs_i, pls); print("python total ms", (time.perf_counter() - t) * 1e3, file=sys.stderr)
t = time.perf_counter(); compositor.composite(bg_i, objs_i, pls); print("python total ms", (time.perf_counter() - t) * 1e3, file=sys.stderr)
