"""Debug: run a golden case through CompositeBatch and print its command streams (B200COMP_DBG=1 walks
the streams without issuing copies, so the dump survives a faulting kernel)."""
import ctypes, sys
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch
import golden_io as G
from image_transformation_b200 import _native
from image_transformation_b200.batch import CutoutPool, CompositeBatch

name = sys.argv[1] if len(sys.argv) > 1 else "c1_squarespace_1x1"
bg, objs, pl, exp = G.case(name)
pool = CutoutPool({int(k): v for k, v in objs.items()})
bgt = torch.from_numpy(bg).cuda()
b = CompositeBatch(pool, [(bg.shape[1], bg.shape[0])], [pl], backgrounds=[bgt])
print("info", b.info)
b.run(); 
try:
    b.check()
    out = b.output(0).cpu().numpy()
    print("mismatching pixels", int((out != exp).any(axis=2).sum()))
except Exception as e:
    print("run failed:", e)
    sys.exit(1)
lib = _native.lib()
lib.b200comp_plan_debug_streams_.restype = ctypes.c_int64
lib.b200comp_plan_debug_streams_.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
cap = 200000
recs = np.zeros((cap, 16), np.uint32); offs = np.zeros(2048, np.int64); ns = ctypes.c_int(0)
n = lib.b200comp_plan_debug_streams_(b._plan, recs.ctypes.data, cap, offs.ctypes.data, ctypes.byref(ns))
print("records", n, "streams", ns.value)
KIND = {0: "TILE", 1: "RESAMPLE", 2: "IDENT_TMA", 3: "IDENT_LDG", 4: "NOP", 5: "END"}
shown = 0
for c in range(ns.value):
    lo = int(offs[c]); hi = lo
    while hi < n and recs[hi][0] != 5: hi += 1
    hi += 1  # streams are stored in allocation order, each ends with END
    if hi - lo <= 1: continue
    if shown >= int(sys.argv[2]) if len(sys.argv) > 2 else shown >= 12: break
    shown += 1
    print(f"stream {c}: [{lo},{hi})")
    for r in recs[lo:hi]:
        k = int(r[0])
        if k == 0:
            print(f"   TILE steps={r[1]} tx0={r[2]} ty0={r[3]} tw={r[4] & 0xffff} th={r[4] >> 16} flags={r[6]} canvas={r[7]} bg_map={int(r[8]) | int(r[9]) << 32:#x} out_map={int(r[10]) | int(r[11]) << 32:#x}")
        elif k == 1:
            print(f"   RESAMPLE nwx={r[1]&255} nwy={(r[1]>>8)&255} nch={(r[1]>>16)&255} NRQ={r[1]>>24} dx={r[2]&255} dy={(r[2]>>8)&255} two={(r[2]>>16)&255} tho={r[2]>>24} ox0={r[3]} oy0={r[4]} cw0={r[5]&0xffff} rw0={r[5]>>16} w={r[6]} h={r[7]} map={r[8]} plx={r[9]} ply={r[10]} pbw={r[11]&0xffff} nrbox={r[11]>>16}")
        elif k in (2, 3):
            print(f"   {KIND[k]} dx={r[2]&255} dy={(r[2]>>8)&255} two={(r[2]>>16)&255} tho={r[2]>>24} cx={np.int32(r[3])} cy={np.int32(r[4])} map={r[8]} pitch={r[6]}")
        else:
            print("  ", KIND.get(k, f"?{k}"))
