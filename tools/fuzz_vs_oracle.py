"""Randomised differential test of the batched path against the CPU oracle (beyond the fixed seeds of tests/):
random canvas sizes, cutout kinds (opaque rectangles, soft masks, binary alpha, fully transparent, tiny), scales
0.3..3 with single-axis and identity cases, boxes hanging off the canvas, solid / opaque / translucent backgrounds.
usage (GPU box): python tools/fuzz_vs_oracle.py [iterations] [seed]"""
import sys

sys.path.insert(0, ".")
import numpy as np
import torch

import oracle
from image_transformation_b200 import synth
from image_transformation_b200.batch import CompositeBatch, CutoutPool


def run(iters: int, seed: int) -> int:
    rng = np.random.default_rng(seed)
    bad = 0
    for it in range(iters):
        pool = {}
        for k in range(1, int(rng.integers(3, 9))):
            sw, sh = int(rng.integers(1, 420)), int(rng.integers(1, 420))
            kind = rng.integers(0, 5)
            if kind == 0:
                a = rng.integers(0, 256, (sh, sw, 4), dtype=np.uint8); a[..., 3] = 255
            elif kind == 1:
                a = synth.make_cutout(rng, max(sw, 8), max(sh, 8))
            elif kind == 2:
                a = rng.integers(0, 256, (sh, sw, 4), dtype=np.uint8); a[..., 3] = np.where(rng.random((sh, sw)) < 0.5, 0, 255)
            elif kind == 3:
                a = rng.integers(0, 256, (sh, sw, 4), dtype=np.uint8); a[..., 3] = 0
            else:
                a = rng.integers(0, 256, (sh, sw, 4), dtype=np.uint8)
            pool[k] = a
        n_canv = int(rng.integers(1, 4))
        canv, pls, bgs, solids = [], [], [], []
        for c in range(n_canv):
            W, H = int(rng.integers(40, 1100)), int(rng.integers(40, 900))
            canv.append((W, H))
            pl = []
            for _ in range(int(rng.integers(1, 45))):
                oid = int(rng.integers(1, len(pool) + 1))
                sh, sw = pool[oid].shape[:2]
                m = rng.random()
                if m < 0.25: w, h = sw, sh
                elif m < 0.35: w, h = sw, max(1, int(sh * rng.uniform(0.3, 3)))
                elif m < 0.45: w, h = max(1, int(sw * rng.uniform(0.3, 3))), sh
                else:
                    s = rng.uniform(0.3, 3.0); w, h = max(1, int(sw * s)), max(1, int(sh * s * rng.uniform(0.8, 1.25)))
                x, y = int(rng.integers(-w, W)), int(rng.integers(-h, H))
                pl.append({"object_id": oid, "box": [x, y, x + w, y + h]})
            pls.append(pl)
            b = rng.integers(0, 3)
            if b == 0:
                bgs.append(None); solids.append(tuple(int(v) for v in rng.integers(0, 256, 4)))
            else:
                bg = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
                if b == 1: bg[..., 3] = 255
                bgs.append(bg); solids.append((0, 0, 0, 0))
        dp = CutoutPool(pool)
        cb = CompositeBatch(dp, canv, pls, backgrounds=[None if b is None else torch.from_numpy(b).cuda() for b in bgs], solid=solids)
        cb.run(); cb.check()
        for c in range(n_canv):
            W, H = canv[c]
            bg = bgs[c]
            if bg is None:
                bg = np.empty((H, W, 4), np.uint8); bg[...] = solids[c]
            exp = oracle.composite(bg, pool, pls[c])
            got = cb.output(c).cpu().numpy()
            if not np.array_equal(got, exp):
                bad += 1
                d = np.argwhere((got != exp).any(-1))
                print(f"MISMATCH iter {it} canvas {c} {W}x{H}: {len(d)} pixels, first {d[:3].tolist()}", flush=True)
        cb.close()
    return bad


if __name__ == "__main__":
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    bad = run(iters, seed)
    print(f"fuzz seed {seed}: {iters} iterations, {bad} mismatching canvases")
    sys.exit(1 if bad else 0)
