"""Drop-in for ``/root/reference/background_resizing.py``: background colour statistics
and solid / gradient canvas synthesis on the GPU (histogram-median kernel + vectorised
fill), same function names, arguments and return types.
"""
from __future__ import annotations

import ctypes
from typing import Tuple

import numpy as np
from PIL import Image

from . import _native

RGB = Tuple[int, int, int]


def _load_background_rgba(background_path: str) -> Image.Image:
    """background_resizing.py:6-8 (PNG decode stays on the host)."""
    return Image.open(background_path).convert("RGBA")


def _as_rgba_array(img: Image.Image) -> np.ndarray:
    return _native.rgba_array(img if img.mode == "RGBA" else img.convert("RGBA"))


def _median_color_nontransparent(img_rgba: Image.Image) -> RGB:
    """Per-channel median over alpha>0 pixels, all pixels if none (background_resizing.py:11-22)."""
    _native.require_gpu()
    a = _native.rgba_array(img_rgba) if img_rgba.mode == "RGBA" else np.ascontiguousarray(np.array(img_rgba), dtype=np.uint8)
    if a.ndim != 3 or a.shape[2] != 4:
        # the reference indexes arr[:, :, 3]; anything but 4 channels fails there
        raise IndexError("index 3 is out of bounds for axis 2 with size %d" % (a.shape[2] if a.ndim == 3 else 0))
    H, W = a.shape[:2]
    out = (ctypes.c_int32 * 3)()
    rc = _native.lib().b200comp_masked_median_rgb_host(a.ctypes.data, W, H, a.strides[0], 0, 0, W, H, out)
    _native.check(rc, "_median_color_nontransparent")
    return int(out[0]), int(out[1]), int(out[2])


def fill_solid(background_path: str, canvas_size: Tuple[int, int]) -> Image.Image:
    """Solid RGBA canvas in the median non-transparent colour of background.png
    (background_resizing.py:25-33)."""
    bg = _load_background_rgba(background_path)
    W, H = (int(canvas_size[0]), int(canvas_size[1]))
    if W < 1 or H < 1:
        # Image.new accepts a zero-sized canvas; nothing to synthesise on the device then
        color = _median_color_nontransparent(bg)
        return Image.new("RGBA", (W, H), color + (255,))
    _native.require_gpu()
    a = _as_rgba_array(bg)
    result, out = _native.new_rgba_image(W, H)  # the library writes into an image Pillow owns (fully mutable)
    if result is None:
        out = np.empty((H, W, 4), np.uint8)
    rgb = (ctypes.c_int32 * 3)()
    rc = _native.lib().b200comp_fill_solid_host(a.ctypes.data, a.shape[1], a.shape[0], a.strides[0], out.ctypes.data,
                                                W, H, out.strides[0], rgb)
    _native.check(rc, "fill_solid")
    img = result if result is not None else _native.image_from_rgba(out)
    # composite() passes this colour as a value instead of uploading W*H*4 bytes, as long as the canvas is still
    # all this colour when it gets there (compositor._solid_colour_of)
    img._b200_solid = int(rgb[0]) | int(rgb[1]) << 8 | int(rgb[2]) << 16 | 255 << 24
    return img


def _edge_strip_median_colors(img: Image.Image, strip_px: int = 8) -> Tuple[RGB, RGB, RGB, RGB]:
    """Medians of the left / right / top / bottom edge strips (background_resizing.py:36-55)."""
    _native.require_gpu()
    a = _as_rgba_array(img)
    H, W = a.shape[:2]
    out = (ctypes.c_int32 * 12)()
    rc = _native.lib().b200comp_edge_strip_medians_host(a.ctypes.data, W, H, a.strides[0], int(strip_px), out)
    _native.check(rc, "_edge_strip_median_colors")
    v = [int(x) for x in out]
    return tuple(v[0:3]), tuple(v[3:6]), tuple(v[6:9]), tuple(v[9:12])  # type: ignore[return-value]


def _axis_variance(c1: RGB, c2: RGB) -> float:
    """Squared colour distance used as the variance proxy (background_resizing.py:58-60)."""
    return float(sum((int(a) - int(b)) ** 2 for a, b in zip(c1[:3], c2[:3])))


def fill_gradient(background_path: str, canvas_size: Tuple[int, int]) -> Image.Image:
    """Linear gradient between the edge-strip medians along the axis with the lower
    colour distance (background_resizing.py:63-98)."""
    bg = _load_background_rgba(background_path)
    W, H = (int(canvas_size[0]), int(canvas_size[1]))
    _native.require_gpu()
    a = _as_rgba_array(bg)
    result, out = _native.new_rgba_image(W, H)
    if result is None:
        out = np.empty((H, W, 4), np.uint8)
    edges = (ctypes.c_int32 * 12)()
    horizontal = ctypes.c_int(0)
    rc = _native.lib().b200comp_fill_gradient_host(a.ctypes.data, a.shape[1], a.shape[0], a.strides[0],
                                                   out.ctypes.data, W, H, out.strides[0], 8, edges,
                                                   ctypes.byref(horizontal))
    _native.check(rc, "fill_gradient")
    return result if result is not None else _native.image_from_rgba(out)
