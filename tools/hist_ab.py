"""usage (GPU box): python tools/hist_ab.py  -- kernel time of the statistics pass on the 8K backgrounds of C5 (events around 20 calls)"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_transformation_b200 import batch as B, synth

W, H = 7680, 4320
noisy = torch.from_numpy(synth.synthetic_background(W, H)).cuda()
flat = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda"); flat[...] = torch.tensor([220, 238, 245, 255], dtype=torch.uint8, device="cuda")
opaque = noisy.clone(); opaque[..., 3] = 255
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, t in (("noisy, 30 % transparent", noisy), ("noisy, opaque", opaque), ("flat", flat)):
    for _ in range(3): B.masked_median_rgb(t)
    ts = []
    for _ in range(20):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = B.masked_median_rgb(t); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"{name:28s} median {statistics.median(ts):7.1f} us per call (3 kernels + memset + D2H), result {r}")
