"""Drop-in for ``/root/reference/compositor.py``: same functions, same arguments,
same error behaviour, PIL in / PIL out -- the raster work runs on the GPU through
``b200comp_composite_host`` (fused resample + alpha-over tile kernel).
"""
from __future__ import annotations

import ctypes
import json
import os
from typing import Dict, List, Sequence, Tuple

import numpy as np
from PIL import Image

from . import _native

# Pillow >= 12 resizes very tall images vertical-first (PIL Image.py:2431-2435); Pillow 11.3
# (the reference's pin, requirements.txt:29) does not.  Parity target = the installed oracle.
TALL_IMAGE_VERTICAL_FIRST = os.environ.get("B200COMP_PILLOW_COMPAT", "12") != "11"


def resolve_placements(placements: Sequence[dict], sizes: Dict[int, Tuple[int, int]]) -> List[Tuple[int, int, int, int, int, int]]:
    """Host-side coercions of compositor.py:12-18 -> [(object_id, x, y, w, h, flags)].

    ``sizes`` maps object id -> (sw, sh).  Unknown ids are skipped before their box is
    looked at; ids and box values go through ``int()`` (truncation toward zero, ValueError /
    TypeError / KeyError propagate exactly as in the reference); w, h are clamped to >= 1.
    """
    out = []
    for p in placements:
        raw = p["object_id"]
        oid = raw if isinstance(raw, int) else int(raw)
        if oid not in sizes:
            continue
        x1, y1, x2, y2 = (int(v) for v in p["box"])
        w = max(1, x2 - x1)
        h = max(1, y2 - y1)
        sw, sh = sizes[oid]
        flags = _native.VERTICAL_FIRST if (TALL_IMAGE_VERTICAL_FIRST and sh > 100 * sw and h < sh) else 0
        out.append((oid, x1, y1, w, h, flags))
    return out


def _rgba_array(img: Image.Image) -> np.ndarray:
    return _native.rgba_array(img)


def composite(background_img: Image.Image, object_images: Dict[int, Image.Image], placements: List[Dict]) -> Image.Image:
    """Composite objects onto the background according to placements (compositor.py:6-22).

    placements: list of {object_id, box: [x1, y1, x2, y2]}; list order is z-order.
    The background is not modified; a new RGBA image is returned.
    """
    sizes = {oid: im.size for oid, im in object_images.items()}
    resolved = resolve_placements(placements, sizes)
    if not resolved:
        return background_img.copy()
    # Image.alpha_composite's own checks, in the order the reference would hit them
    if background_img.mode != "RGBA":
        raise ValueError("image has wrong mode")
    for oid, *_ in resolved:
        if object_images[oid].mode != "RGBA":
            raise ValueError("images do not match")

    _native.require_gpu()
    W, H = background_img.size
    bg = _rgba_array(background_img)
    out = np.empty((H, W, 4), np.uint8)
    arrays: Dict[int, np.ndarray] = {}
    recs = (_native.Placement * len(resolved))()
    for i, (oid, x, y, w, h, flags) in enumerate(resolved):
        if oid not in arrays:
            arrays[oid] = _rgba_array(object_images[oid])
        a = arrays[oid]
        recs[i] = _native.Placement(a.ctypes.data, a.strides[0], a.shape[1], a.shape[0], x, y, w, h, flags, 0)
    rc = _native.lib().b200comp_composite_host(bg.ctypes.data, W, H, bg.strides[0], out.ctypes.data, out.strides[0],
                                               recs, len(resolved))
    _native.check(rc, "composite")
    return _native.image_from_rgba(out)


def load_object_images(results_json_path: str) -> Dict[int, Image.Image]:
    """{object_id: RGBA cutout} from a bundle's results.json (compositor.py:25-35). Host I/O."""
    with open(results_json_path, "r", encoding="utf-8") as f:
        items = json.load(f)
    root = os.path.dirname(results_json_path)
    return {int(it["object_id"]): Image.open(os.path.join(root, it["filename"])).convert("RGBA") for it in items}
