#!/usr/bin/env python3
"""Achieved HBM GB/s of the stand-alone kernels around the fused path (north_star kernels 1 and 3): background
statistics (masked histogram median), solid and gradient fill, stand-alone LANCZOS resize, stand-alone alpha-over.
Algorithmic bytes as in SURVEY.md 8(d); CUDA events on the launching stream, L2 flushed between timed launches
(a 256 MB write), median of 20.

    python profiles/aux_kernels_bench.py > profiles/r1_aux_kernels.json      (on the GPU box)
"""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from image_transformation_b200 import batch as B, synth  # noqa: E402


def timed(fn, reps=20):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def main():
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    W, H = 7680, 4320  # C5 canvas / background size
    rows = []

    def row(name, algo_bytes, ms, note=""):
        gbs = algo_bytes / 1e9 / (ms / 1e3)
        rows.append({"kernel": name, "algorithmic_bytes": int(algo_bytes), "ms": ms, "GBps": gbs, "frac_of_peak": gbs / peak, "note": note})

    bg = torch.from_numpy(synth.synthetic_background(W, H)).cuda()
    ms = timed(lambda: B.masked_median_rgb(bg))
    row("hist_rgb_kernel + median_from_hist_kernel (fill_solid statistics, 8K background)", 4 * W * H, ms,
        "includes the D2H of the 3 medians and the host sync of the call")
    flat = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda")
    flat[...] = torch.tensor([220, 238, 245, 255], dtype=torch.uint8, device="cuda")
    ms = timed(lambda: B.masked_median_rgb(flat))
    row("hist_rgb_kernel + median_from_hist_kernel (flat 8K background: every pixel in the same three bins)", 4 * W * H, ms,
        "includes the D2H of the 3 medians and the host sync of the call")
    del flat
    canvas = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda")
    ms = timed(lambda: B.fill_rgba_(canvas, (38, 73, 115, 255)))
    row("fill_flat_kernel (solid 8K canvas)", 4 * W * H, ms)
    ms = timed(lambda: B.fill_gradient_(canvas, True, (10, 20, 30), (200, 180, 90)))
    row("gradient fill (8K canvas)", 4 * W * H, ms)
    rng = np.random.default_rng(1)
    src = torch.from_numpy(synth.make_cutout(rng, 3072, 2304)).cuda()
    for (w, h) in ((2304, 1728), (1536, 1152), (4096, 3072), (768, 576), (384, 288)):
        ms = timed(lambda: B.resize_rgba_lanczos(src, (w, h)))
        path = "fused tile kernel, replace mode" if 3072 / w <= 2.6 else "generic two-pass kernels: more than 17 taps"
        row(f"resize_rgba_lanczos 3072x2304 -> {w}x{h} ({path})", 4 * 3072 * 2304 + 4 * w * h, ms,
            "stand-alone call: plan creation for one placement included")
    ov = torch.from_numpy(synth.make_cutout(rng, 3072, 2304)).cuda()
    big = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda")
    B.fill_rgba_(big, (1, 2, 3, 255))
    ms = timed(lambda: B.alpha_over_(big, ov, (1000, 700)))
    row("alpha_over_kernel (3072x2304 overlay onto an 8K canvas)", 3 * 4 * 3072 * 2304, ms)
    print(json.dumps({"peak_GBps": peak, "device": torch.cuda.get_device_name(0), "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
