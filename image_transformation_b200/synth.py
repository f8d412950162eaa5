"""Synthetic bundles and layouts for the benchmark configurations of BASELINE.json
(SURVEY.md section 8d, C3-C5).  Host-side NumPy only; produces exactly what the reference's
``composite()`` consumes: {object_id: RGBA array} and per-canvas placement lists
({object_id, box}), so the same inputs feed the CUDA path, the oracle and the CPU baseline.

The layout resolver below is a small flex-style placer written for this generator (rows /
columns, justify, align, gap, then the clamp-into-canvas rule of
macro_placement_test.py:954-964).  Flex-DSL itself stays in the reference's host Python and
is out of scope; the hot path only ever sees the resulting integer boxes.  Workload
``c3_refplacer`` uses layouts resolved by the reference's own placer instead (a committed fixture):
its coverage statistics agree with this generator's.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

Placement = Dict[str, object]


# ---------------------------------------------------------------------------- cutouts
def make_cutout(rng: np.random.Generator, sw: int, sh: int) -> np.ndarray:
    """Random-colour RGBA cutout with a soft-edged ellipse or rounded-rectangle alpha mask
    (roughly 30 % transparent, 60 % opaque, 10 % partial alpha)."""
    a = np.empty((sh, sw, 4), np.uint8)
    a[..., :3] = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:sh, 0:sw].astype(np.float32)
    cx, cy = (sw - 1) / 2.0, (sh - 1) / 2.0
    edge = 0.055  # soft band, in units of the half-size
    if rng.random() < 0.5:
        r = np.sqrt(((xx - cx) / (sw / 2.0)) ** 2 + ((yy - cy) / (sh / 2.0)) ** 2)
        d = (0.90 - r) / edge
    else:
        # rounded rectangle inset by a margin
        m, rad = 0.16, 0.35
        ux = np.abs(xx - cx) / (sw / 2.0)
        uy = np.abs(yy - cy) / (sh / 2.0)
        qx = np.maximum(ux - (1 - m - rad), 0)
        qy = np.maximum(uy - (1 - m - rad), 0)
        d = (rad - np.sqrt(qx * qx + qy * qy)) / edge
    a[..., 3] = np.clip(d * 255.0 + 127.0, 0, 255).astype(np.uint8)
    return a


def make_pool(n: int = 64, lo: int = 256, hi: int = 1536, seed: int = 1234) -> Dict[int, np.ndarray]:
    rng = np.random.default_rng(seed)
    pool = {}
    for oid in range(1, n + 1):
        sw, sh = (int(v) for v in rng.integers(lo, hi + 1, 2))
        pool[oid] = make_cutout(rng, sw, sh)
    return pool


def alpha_stats(pool: Dict[int, np.ndarray]) -> Dict[str, float]:
    tot = sum(a.shape[0] * a.shape[1] for a in pool.values())
    z = sum(int((a[..., 3] == 0).sum()) for a in pool.values())
    o = sum(int((a[..., 3] == 255).sum()) for a in pool.values())
    return {"transparent": z / tot, "opaque": o / tot, "partial": 1 - (z + o) / tot}


# ---------------------------------------------------------------------------- layouts
def _justify(sizes: Sequence[int], extent: int, gap: int, mode: str) -> List[int]:
    n = len(sizes)
    used = sum(sizes) + gap * max(0, n - 1)
    free = extent - used
    if mode == "center":
        pos, step = free / 2.0, gap
    elif mode == "end":
        pos, step = free, gap
    elif mode == "space-between" and n > 1:
        pos, step = 0.0, gap + max(0.0, free) / (n - 1)
    elif mode == "space-around":
        pad = max(0.0, free) / (2 * n) if n else 0
        pos, step = pad, gap + 2 * pad
    else:
        pos, step = 0.0, gap
    out = []
    for s in sizes:
        out.append(int(round(pos)))
        pos += s + step
    return out


def _align(size: int, extent: int, mode: str) -> int:
    if mode == "center":
        return int(round((extent - size) / 2.0))
    if mode == "end":
        return extent - size
    return 0


def flex_layout(rng: np.random.Generator, canvas: Tuple[int, int], items: Sequence[Tuple[int, int, int]]) -> List[Placement]:
    """Random two-level row/column tree over `items` = [(object_id, w, h)]; returns boxes
    clamped into the canvas the way _clamp_boxes_to_canvas does (shift inside when the box
    fits, otherwise pin to 0 and let it overhang)."""
    W, H = canvas
    modes = ("start", "center", "end", "space-between", "space-around")
    aligns = ("start", "center", "end")
    direction = "row" if rng.random() < 0.5 else "column"
    n_groups = int(rng.integers(2, 6))
    order = list(items)
    groups: List[List[Tuple[int, int, int]]] = [[] for _ in range(n_groups)]
    for i, it in enumerate(order):
        groups[i % n_groups].append(it)
    groups = [g for g in groups if g]
    gap = int(rng.integers(0, 41))
    jmode, amode = modes[int(rng.integers(0, 5))], aligns[int(rng.integers(0, 3))]
    sub = []
    for g in groups:
        sdir = "column" if direction == "row" else "row"
        sgap = int(rng.integers(0, 41))
        if sdir == "row":
            gw = sum(w for _, w, _ in g) + sgap * (len(g) - 1)
            gh = max(h for _, _, h in g)
        else:
            gw = max(w for _, w, _ in g)
            gh = sum(h for _, _, h in g) + sgap * (len(g) - 1)
        sub.append((g, sdir, sgap, gw, gh, modes[int(rng.integers(0, 5))], aligns[int(rng.integers(0, 3))]))
    main = [s[3] if direction == "row" else s[4] for s in sub]
    starts = _justify(main, W if direction == "row" else H, gap, jmode)
    out: List[Placement] = []
    for (g, sdir, sgap, gw, gh, sj, sa), st in zip(sub, starts):
        if direction == "row":
            gx, gy = st, _align(gh, H, amode)
        else:
            gx, gy = _align(gw, W, amode), st
        inner = _justify([w if sdir == "row" else h for _, w, h in g], gw if sdir == "row" else gh, sgap, sj)
        for (oid, w, h), p in zip(g, inner):
            if sdir == "row":
                x, y = gx + p, gy + _align(h, gh, sa)
            else:
                x, y = gx + _align(w, gw, sa), gy + p
            x = max(0, min(x, W - w))
            y = max(0, min(y, H - h))
            out.append({"object_id": oid, "box": [int(x), int(y), int(x + w), int(y + h)]})
    # z-order = placement list order; shuffle so overlap order is not tied to the tree walk
    perm = rng.permutation(len(out))
    return [out[int(i)] for i in perm]


_REF_LAYOUTS = None


def reference_placer_layout(canvas: Tuple[int, int], canvas_idx: int) -> List[Placement]:
    """Layout `canvas_idx` (modulo the fixture's 256) of tests/golden/c3_reference_layouts.npz: the same object draws as
    canvas_placements(), a random Flex-DSL tree over them resolved by the REFERENCE's `_place_flex_container` +
    `_clamp_boxes_to_canvas` on size proxies (generated by tests/golden/make_c3_reference_layouts.py where the reference
    is available).  Coverage statistics match flex_layout's (74 % of the canvas covered, mean depth 1.24)."""
    global _REF_LAYOUTS
    if _REF_LAYOUTS is None:
        import os

        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "c3_reference_layouts.npz")
        d = np.load(path)
        _REF_LAYOUTS = (d["oid"], d["box"], tuple(int(v) for v in d["canvas"]))
    oid, box, size = _REF_LAYOUTS
    if tuple(canvas) != size:
        raise ValueError(f"reference-placer layouts exist for {size} canvases only")
    i = canvas_idx % oid.shape[0]
    return [{"object_id": int(o), "box": [int(v) for v in b]} for o, b in zip(oid[i], box[i]) if o > 0]


def canvas_placements(pool_sizes: Dict[int, Tuple[int, int]], canvas: Tuple[int, int], canvas_idx: int,
                      n_objects: int = 20, seed: int = 1234, scale_lo: float = 0.5, scale_hi: float = 1.0,
                      identity_frac: float = 0.1, layout: str = "flex") -> List[Placement]:
    """Placements of one synthetic canvas: `n_objects` draws from the pool, isotropic scale
    (exactly 1.0 with probability `identity_frac`, else U[scale_lo, scale_hi])."""
    rng = np.random.default_rng(seed + canvas_idx)
    ids = sorted(pool_sizes)
    items = []
    for _ in range(n_objects):
        oid = ids[int(rng.integers(0, len(ids)))]
        sw, sh = pool_sizes[oid]
        s = 1.0 if rng.random() < identity_frac else float(rng.uniform(scale_lo, scale_hi))
        items.append((oid, max(1, int(round(sw * s))), max(1, int(round(sh * s)))))
    if layout == "refplacer":
        return reference_placer_layout(canvas, canvas_idx)
    if layout == "flex":
        return flex_layout(rng, canvas, items)
    W, H = canvas
    out = []
    for oid, w, h in items:  # uniform positions (heavy overlap), partly overhanging
        x = int(rng.integers(-w // 4, max(1, W - (3 * w) // 4)))
        y = int(rng.integers(-h // 4, max(1, H - (3 * h) // 4)))
        out.append({"object_id": oid, "box": [x, y, x + w, y + h]})
    return out


# ---------------------------------------------------------------------------- named workloads
WORKLOADS = {
    # BASELINE.json configs[2]: synthetic 4K canvases, 20 RGBA objects each, batch 1024
    "c3_4k_20obj": dict(canvas=(3840, 2160), n_objects=20, pool_n=64, pool_lo=256, pool_hi=1536,
                        scale_lo=0.5, scale_hi=1.0, layout="flex", batch=1024),
    # C3 with the layouts resolved by the reference's own Flex-DSL placer (fixture, 256 layouts repeated)
    "c3_refplacer": dict(canvas=(3840, 2160), n_objects=20, pool_n=64, pool_lo=256, pool_hi=1536,
                         scale_lo=0.5, scale_hi=1.0, layout="refplacer", batch=1024),
    # kernel-tuning variant of C3: scales 0.75..1.0 only (smaller source patches, 9-tap windows)
    "c3_s75": dict(canvas=(3840, 2160), n_objects=20, pool_n=64, pool_lo=256, pool_hi=1536,
                   scale_lo=0.75, scale_hi=1.0, layout="flex", batch=1024),
    # configs[3]: aspect sweep at the same 8.29 MP budget
    "c4_aspect_sweep": dict(canvases=[(2160, 3840), (2880, 2880), (3840, 2160), (4399, 1885)], n_objects=20,
                            pool_n=64, pool_lo=256, pool_hi=1536, scale_lo=0.5, scale_hi=1.0, layout="flex",
                            batch=4096),
    # configs[4]: 8K canvases, 64 large overlapping objects
    "c5_8k_64obj": dict(canvas=(7680, 4320), n_objects=64, pool_n=24, pool_lo=1024, pool_hi=3072,
                        scale_lo=0.6, scale_hi=1.0, layout="uniform", batch=64),
}


def workload_canvas_size(name: str, canvas_idx: int) -> Tuple[int, int]:
    w = WORKLOADS[name]
    if "canvases" in w:
        return w["canvases"][canvas_idx % len(w["canvases"])]
    return w["canvas"]


def workload_pool(name: str, seed: int = 1234) -> Dict[int, np.ndarray]:
    w = WORKLOADS[name]
    return make_pool(w["pool_n"], w["pool_lo"], w["pool_hi"], seed)


def workload_placements(name: str, pool_sizes: Dict[int, Tuple[int, int]], canvas_idx: int, seed: int = 1234) -> List[Placement]:
    w = WORKLOADS[name]
    return canvas_placements(pool_sizes, workload_canvas_size(name, canvas_idx), canvas_idx, w["n_objects"], seed,
                             w["scale_lo"], w["scale_hi"], 0.1, w["layout"])


def synthetic_background(W: int, H: int, seed: int = 99) -> np.ndarray:
    """Smooth RGB + noise with binary alpha (about 30 % zero), like the bundles' background.png."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    a = np.empty((H, W, 4), np.uint8)
    for c, (fx, fy) in enumerate(((0.7, 0.2), (0.3, 0.9), (0.5, 0.5))):
        base = 60 + 120 * (fx * xx / max(1, W - 1) + fy * yy / max(1, H - 1)) / (fx + fy)
        a[..., c] = np.clip(base + rng.normal(0, 6, (H, W)), 0, 255).astype(np.uint8)
    holes = (np.sin(xx / (W / 9.0)) * np.cos(yy / (H / 5.0))) > 0.42
    a[..., 3] = np.where(holes, 0, 255).astype(np.uint8)
    return a
