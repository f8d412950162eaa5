"""Race hunting without a sanitizer: the same culling-heavy batches run many times (different CUDA-side timing
every run) must give byte-identical canvases, and the first canvas must match the oracle.
usage (GPU box): python tools/stress_determinism.py [runs]"""
import hashlib
import sys

sys.path.insert(0, ".")
import numpy as np
import torch

import oracle
from image_transformation_b200 import synth
from image_transformation_b200.batch import CompositeBatch, CutoutPool

runs = int(sys.argv[1]) if len(sys.argv) > 1 else 12
for name, n, check in (("c5_8k_64obj", 3, True), ("c3_4k_20obj", 24, False), ("c4_aspect_sweep", 8, False)):
    pool_np = synth.workload_pool(name)
    sizes = {k: (v.shape[1], v.shape[0]) for k, v in pool_np.items()}
    pool = CutoutPool(pool_np)
    canv = [synth.workload_canvas_size(name, i) for i in range(n)]
    pls = [synth.workload_placements(name, sizes, i) for i in range(n)]
    rng = np.random.default_rng(3)
    bgs = [torch.from_numpy(rng.integers(0, 256, (h, w, 4), dtype=np.uint8)).cuda() for (w, h) in canv]
    cb = CompositeBatch(pool, canv, pls, backgrounds=bgs)
    ref = None
    for r in range(runs):
        cb.run()
        cb.check()
        hs = [hashlib.sha256(o.cpu().numpy().tobytes()).hexdigest() for o in cb.outputs()]
        if ref is None:
            ref = hs
            if check:
                exp = oracle.composite(bgs[0].cpu().numpy(), pool_np, pls[0])
                assert np.array_equal(cb.output(0).cpu().numpy(), exp), f"{name}: canvas 0 differs from the oracle"
        assert hs == ref, f"{name}: run {r} differs from run 0 on canvases {[i for i, (a, b) in enumerate(zip(hs, ref)) if a != b]}"
    print(f"{name}: {runs} runs x {n} canvases identical" + (", canvas 0 == oracle" if check else ""), flush=True)
    cb.close()
