#!/usr/bin/env python3
"""C3 layouts resolved by the REFERENCE's own Flex-DSL placer.

Run in the build container only (needs /root/reference):

    python tests/golden/make_c3_reference_layouts.py

For canvas i (rng seeded 1234 + i, exactly the draws of image_transformation_b200.synth.canvas_placements): 20 objects
from the 64-cutout pool, isotropic scale 1.0 with probability 0.1 else U[0.5, 1]; a random Flex-DSL tree of depth <= 2
(random row / column, justify, align, gap 0-40) over them; resolved by the unmodified
macro_placement_test._place_flex_container on size-proxy objects (the placer only reads .size) and clamped by
_clamp_boxes_to_canvas; z-order = placement list order.  Output: tests/golden/c3_reference_layouts.npz with
  oid   int32 [N, 20]      pool object id of each placement (0 = placement dropped by the placer)
  box   int32 [N, 20, 4]   x1, y1, x2, y2
The fixture travels to the GPU box; bench.py --workload c3_refplacer and tests/test_gpu_parity.py read it.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

with contextlib.redirect_stdout(io.StringIO()):
    import macro_placement_test as ref_mpt  # noqa: E402

from image_transformation_b200 import synth  # noqa: E402

N_CANVASES = 256
CANVAS = (3840, 2160)
JUSTIFY = ("start", "center", "end", "space-between", "space-around")
ALIGN = ("start", "center", "end")


class SizeProxy:
    """The placer only reads .size (macro_placement_test.py:655,709)."""

    def __init__(self, size):
        self.size = size


def random_tree(rng, leaf_ids):
    """Depth <= 2: a root container whose children are leaves or one-level containers of leaves."""
    direction = "row" if rng.random() < 0.5 else "column"
    n_groups = int(rng.integers(2, 6))
    groups = [[] for _ in range(n_groups)]
    for k, leaf in enumerate(leaf_ids):
        groups[k % n_groups].append(leaf)

    def container(d, children):
        return {"direction": d, "justify": JUSTIFY[int(rng.integers(0, 5))], "align": ALIGN[int(rng.integers(0, 3))],
                "gap_px": int(rng.integers(0, 41)), "padding_px": 0, "children": children}

    children = []
    for g in groups:
        if not g:
            continue
        if len(g) == 1:
            children.append({"object_id": g[0]})
        else:
            children.append(container("column" if direction == "row" else "row", [{"object_id": leaf} for leaf in g]))
    return container(direction, children)


def main():
    pool = synth.workload_pool("c3_4k_20obj")
    sizes = {k: (v.shape[1], v.shape[0]) for k, v in pool.items()}
    ids = sorted(sizes)
    oid = np.zeros((N_CANVASES, 20), np.int32)
    box = np.zeros((N_CANVASES, 20, 4), np.int32)
    dropped = 0
    for i in range(N_CANVASES):
        rng = np.random.default_rng(1234 + i)
        items = []
        for _ in range(20):  # the draws of synth.canvas_placements
            o = ids[int(rng.integers(0, len(ids)))]
            sw, sh = sizes[o]
            s = 1.0 if rng.random() < 0.1 else float(rng.uniform(0.5, 1.0))
            items.append((o, max(1, int(round(sw * s))), max(1, int(round(sh * s)))))
        leaves = list(range(1, 21))  # one proxy per placement: the same cutout may appear twice at different scales
        proxies = {leaf: SizeProxy((items[leaf - 1][1], items[leaf - 1][2])) for leaf in leaves}
        tree = random_tree(rng, leaves)
        placements = []
        with contextlib.redirect_stdout(io.StringIO()):
            ref_mpt._place_flex_container(tree, (0, 0), CANVAS, proxies, placements, "root")
            ref_mpt._clamp_boxes_to_canvas(placements, CANVAS)
        perm = rng.permutation(len(placements))  # z-order not tied to the tree walk
        for k, j in enumerate(perm):
            p = placements[int(j)]
            leaf = int(p["object_id"])
            oid[i, k] = items[leaf - 1][0]
            box[i, k] = [int(v) for v in p["box"]]
        dropped += 20 - len(placements)
    out = os.path.join(HERE, "c3_reference_layouts.npz")
    np.savez_compressed(out, oid=oid, box=box, canvas=np.array(CANVAS, np.int32))
    w = box[..., 2] - box[..., 0]
    h = box[..., 3] - box[..., 1]
    print(f"{out}: {N_CANVASES} canvases, {dropped} placements dropped by the placer, box sizes {w[oid > 0].min()}..{w.max()} x "
          f"{h[oid > 0].min()}..{h.max()}, {os.path.getsize(out)} bytes")


if __name__ == "__main__":
    main()
