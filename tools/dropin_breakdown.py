"""Where a small drop-in composite() call spends its time (PIL boundary in Python vs the C call)."""
import os, sys, time, statistics, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from PIL import Image
import golden_io as G
from image_transformation_b200 import compositor as C, _native

bg, objs, pls, exp = G.case("c1_squarespace_1x1")
bg_i = Image.fromarray(bg, "RGBA").copy()
objs_i = {k: Image.fromarray(v, "RGBA").copy() for k, v in objs.items()}
C.composite(bg_i, objs_i, pls)

def med(fn, n=200):
    fn(); ts = []
    for _ in range(n):
        t = time.perf_counter(); fn(); ts.append((time.perf_counter() - t) * 1e6)
    return statistics.median(ts)

sizes = {k: im.size for k, im in objs_i.items()}
print("whole composite()            %8.1f us" % med(lambda: C.composite(bg_i, objs_i, pls)))
print("resolve_placements           %8.1f us" % med(lambda: C.resolve_placements(pls, sizes)))
print("rgba_array(bg)               %8.1f us" % med(lambda: _native.rgba_array(bg_i)))
print("rgba_array x4 cutouts        %8.1f us" % med(lambda: [_native.rgba_array(v) for v in objs_i.values()]))
print("new_rgba_image               %8.1f us" % med(lambda: _native.new_rgba_image(492, 492)))
res = C.resolve_placements(pls, sizes)
arrs = {k: _native.rgba_array(v) for k, v in objs_i.items()}
bga = _native.rgba_array(bg_i)
img, out = _native.new_rgba_image(492, 492)
recs = (_native.Placement * len(res))()
for i, (oid, x, y, w, h, fl) in enumerate(res):
    a = arrs[oid]
    recs[i] = _native.Placement(a.ctypes.data, a.strides[0], a.shape[1], a.shape[0], x, y, w, h, fl, 0)
L = _native.lib()
call = lambda: L.b200comp_composite_host_ex(bga.ctypes.data, 0, 492, 492, bga.strides[0], out.ctypes.data, out.strides[0], recs, len(res))
print("C call composite_host_ex     %8.1f us" % med(call))
os.environ["B200COMP_TRACE_PLAN"] = "1"
call()
