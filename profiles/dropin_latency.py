#!/usr/bin/env python3
"""Latency of the drop-in calls a user of the reference makes one at a time (PIL in / PIL out), next to the same
operations done by Pillow on the host (the reference's compositor.py:11-21 loop, restated with PIL calls).

    python profiles/dropin_latency.py > profiles/r1_dropin_latency.json      (on the GPU box)
"""
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
from PIL import Image  # noqa: E402

import golden_io as G  # noqa: E402
from image_transformation_b200 import compositor, synth  # noqa: E402


def pil_composite(bg, objs, placements):
    canvas = bg.copy()
    for p in placements:
        oid = int(p["object_id"])
        if oid not in objs:
            continue
        x1, y1, x2, y2 = [int(v) for v in p["box"]]
        canvas.alpha_composite(objs[oid].resize((max(1, x2 - x1), max(1, y2 - y1)), Image.LANCZOS), dest=(x1, y1))
    return canvas


def med(fn, reps):
    fn()
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t) * 1e3)
    return statistics.median(ts)


def main():
    rows = []
    # C1 / C2: the reference bundles on their small canvases (golden cases, all identity sizes) + a scaled variant
    for name in G.case_names():
        c = next(c for c in G.manifest()["cases"] if c["name"] == name)
        if "bundle" not in c:
            continue
        bg, objs, pls, exp = G.case(name)
        bg_i = Image.fromarray(bg, "RGBA").copy()  # images Pillow owns, as Image.open(...).convert("RGBA") returns
        objs_i = {k: Image.fromarray(v, "RGBA").copy() for k, v in objs.items()}
        out = compositor.composite(bg_i, objs_i, pls)
        assert np.array_equal(np.asarray(out), exp), name
        rows.append({"case": name, "canvas": list(bg_i.size), "placements": len(pls),
                     "b200_ms": med(lambda: compositor.composite(bg_i, objs_i, pls), 30),
                     "pillow_ms": med(lambda: pil_composite(bg_i, objs_i, pls), 30)})
    # one C3 canvas (3840x2160, 20 objects)
    pool = synth.workload_pool("c3_4k_20obj")
    sizes = {k: (v.shape[1], v.shape[0]) for k, v in pool.items()}
    pls = synth.workload_placements("c3_4k_20obj", sizes, 0)
    bg_i = Image.new("RGBA", (3840, 2160), (38, 73, 115, 255))
    objs_i = {k: Image.fromarray(v, "RGBA").copy() for k, v in pool.items()}
    a = np.asarray(compositor.composite(bg_i, objs_i, pls))
    b = np.asarray(pil_composite(bg_i, objs_i, pls))
    assert np.array_equal(a, b)
    rows.append({"case": "c3_4k_20obj canvas 0", "canvas": [3840, 2160], "placements": len(pls),
                 "b200_ms": med(lambda: compositor.composite(bg_i, objs_i, pls), 10),
                 "pillow_ms": med(lambda: pil_composite(bg_i, objs_i, pls), 3)})
    # the refine loop of macro_placement_test.py:1679-1699 on the squarespace bundle: reload the bundle, composite
    ref = os.path.join(ROOT, "baseline", "_ref", "output", "squarespace")
    loop = None
    if os.path.isdir(ref):
        bg_i = Image.new("RGBA", (492, 492), (220, 238, 245, 255))
        probe = compositor.load_object_images(os.path.join(ref, "results.json"))
        pls = [{"object_id": k, "box": [20 + 30 * i, 15 + 90 * i, 20 + 30 * i + im.size[0], 15 + 90 * i + im.size[1]]}
               for i, (k, im) in enumerate(sorted(probe.items()))]

        def b200_iter():
            return compositor.composite(bg_i, compositor.load_object_images(os.path.join(ref, "results.json")), pls)

        def pillow_iter():
            with open(os.path.join(ref, "results.json")) as f:
                items = json.load(f)
            objs = {int(it["object_id"]): Image.open(os.path.join(ref, it["filename"])).convert("RGBA") for it in items}
            return pil_composite(bg_i, objs, pls)

        assert np.array_equal(np.asarray(b200_iter()), np.asarray(pillow_iter()))
        loop = {"case": "refine-loop iteration: load_object_images + composite (squarespace, 492x492, identity sizes)",
                "b200_ms": med(b200_iter, 30), "pillow_ms": med(pillow_iter, 30)}
    print(json.dumps({"rows": rows, "refine_loop": loop}, indent=1))


if __name__ == "__main__":
    main()
