"""CPU oracle for the compositor hot path -- TEST INFRASTRUCTURE ONLY.

ctypes binding of ``oracle/compositor_oracle.c`` (a from-scratch restatement of the
Pillow / NumPy arithmetic behind ``/root/reference/compositor.py:6-22`` and
``/root/reference/background_resizing.py:11-98``).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package; the product package ``image_transformation_b200``
never does.  Parity pinning: see ``tests/golden/`` and ``tests/test_oracle.py``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, List, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liborc.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (seconds).  Returns the path of the .so."""
    src = os.path.join(_HERE, "compositor_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB_PATH


class _Placement(ctypes.Structure):
    _fields_ = [
        ("src", ctypes.c_void_p),
        ("sw", ctypes.c_int32),
        ("sh", ctypes.c_int32),
        ("x1", ctypes.c_int32),
        ("y1", ctypes.c_int32),
        ("x2", ctypes.c_int32),
        ("y2", ctypes.c_int32),
    ]


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        vp, ci, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
        L.orc_ksize.argtypes = [ci, ci]
        L.orc_ksize.restype = ci
        L.orc_coeffs.argtypes = [ci, ci, vp, vp]
        L.orc_coeffs.restype = ci
        L.orc_premultiply.argtypes = [vp, vp, sz]
        L.orc_unpremultiply.argtypes = [vp, vp, sz]
        L.orc_resample_h.argtypes = [vp, ci, ci, vp, ci, vp, vp, ci]
        L.orc_resample_v.argtypes = [vp, ci, ci, vp, ci, vp, vp, ci]
        L.orc_resize_rgba_lanczos.argtypes = [vp, ci, ci, vp, ci, ci, ci]
        L.orc_resize_rgba_lanczos.restype = ci
        L.orc_alpha_over_inplace.argtypes = [vp, ci, ci, vp, ci, ci, ci, ci]
        L.orc_composite.argtypes = [vp, ci, ci, vp, ci, ctypes.POINTER(_Placement), ci]
        L.orc_composite.restype = ci
        L.orc_masked_median_rgb.argtypes = [vp, ci, ci, ci, ci, ci, ci, vp]
        L.orc_fill_rgba.argtypes = [vp, ci, ci, ci, ci, ci, ci]
        L.orc_fill_gradient.argtypes = [vp, ci, ci, ci, vp, vp]
        _lib = L
    return _lib


def _u8(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.ndim == 3 and a.shape[2] == 4, "expected HxWx4 uint8 RGBA"
    return a


def _p(a: np.ndarray) -> int:
    return a.ctypes.data


def coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """(k[out_size, ksize] int32, bounds[out_size, 2] int32, ksize)."""
    L = lib()
    ks = L.orc_ksize(in_size, out_size)
    k = np.zeros((out_size, ks), np.int32)
    b = np.zeros((out_size, 2), np.int32)
    L.orc_coeffs(in_size, out_size, _p(k), _p(b))
    return k, b, ks


def premultiply(img: np.ndarray) -> np.ndarray:
    img = _u8(img)
    out = np.empty_like(img)
    lib().orc_premultiply(_p(img), _p(out), img.shape[0] * img.shape[1])
    return out


def unpremultiply(img: np.ndarray) -> np.ndarray:
    img = _u8(img)
    out = np.empty_like(img)
    lib().orc_unpremultiply(_p(img), _p(out), img.shape[0] * img.shape[1])
    return out


def resize_rgba_lanczos(src: np.ndarray, size: Tuple[int, int], vertical_first_rule: bool = True) -> np.ndarray:
    """``Image.resize(size, LANCZOS)`` on an RGBA array; size = (w, h)."""
    src = _u8(src)
    w, h = int(size[0]), int(size[1])
    out = np.empty((h, w, 4), np.uint8)
    rc = lib().orc_resize_rgba_lanczos(_p(src), src.shape[1], src.shape[0], _p(out), w, h, int(vertical_first_rule))
    if rc != 0:
        raise ValueError("oracle resize: bad size")
    return out


def alpha_over_inplace(canvas: np.ndarray, overlay: np.ndarray, dest: Tuple[int, int]) -> None:
    assert canvas.dtype == np.uint8 and canvas.flags.c_contiguous
    overlay = _u8(overlay)
    lib().orc_alpha_over_inplace(
        _p(canvas), canvas.shape[1], canvas.shape[0], _p(overlay), overlay.shape[1], overlay.shape[0],
        int(dest[0]), int(dest[1]),
    )


def composite(background: np.ndarray, objects: Dict[int, np.ndarray], placements: Sequence[dict],
              vertical_first_rule: bool = True) -> np.ndarray:
    """Restates ``composite`` (compositor.py:6-22) on arrays, including the host-side
    id coercion, unknown-id skip, int() truncation and max(1, .) clamps."""
    bg = _u8(background)
    keep: List[np.ndarray] = []
    recs = []
    for p in placements:
        oid = int(p["object_id"]) if not isinstance(p["object_id"], int) else p["object_id"]
        if oid not in objects:
            continue
        x1, y1, x2, y2 = [int(v) for v in p["box"]]
        src = _u8(objects[oid])
        keep.append(src)
        recs.append(_Placement(_p(src), src.shape[1], src.shape[0], x1, y1, x2, y2))
    arr = (_Placement * max(1, len(recs)))(*recs)
    out = np.empty_like(bg)
    rc = lib().orc_composite(_p(bg), bg.shape[1], bg.shape[0], _p(out), len(recs), arr, int(vertical_first_rule))
    if rc != 0:
        raise MemoryError("oracle composite failed")
    return out


def masked_median_rgb(img: np.ndarray, rect: Tuple[int, int, int, int] | None = None) -> Tuple[int, int, int]:
    img = _u8(img)
    H, W = img.shape[:2]
    x0, y0, x1, y1 = rect if rect is not None else (0, 0, W, H)
    out = np.zeros(3, np.int32)
    lib().orc_masked_median_rgb(_p(img), W, H, x0, y0, x1, y1, _p(out))
    return int(out[0]), int(out[1]), int(out[2])


def edge_strip_median_colors(img: np.ndarray, strip_px: int = 8):
    """left, right, top, bottom medians (background_resizing.py:36-55)."""
    img = _u8(img)
    h, w = img.shape[:2]
    return (
        masked_median_rgb(img, (0, 0, min(strip_px, w), h)),
        masked_median_rgb(img, (max(0, w - strip_px), 0, w, h)),
        masked_median_rgb(img, (0, 0, w, min(strip_px, h))),
        masked_median_rgb(img, (0, max(0, h - strip_px), w, h)),
    )


def fill_solid_from(bg_rgba: np.ndarray, canvas_size: Tuple[int, int]) -> np.ndarray:
    """fill_solid (background_resizing.py:25-33) after the PNG decode."""
    r, g, b = masked_median_rgb(bg_rgba)
    W, H = canvas_size
    out = np.empty((H, W, 4), np.uint8)
    lib().orc_fill_rgba(_p(out), W, H, r, g, b, 255)
    return out


def fill_gradient_from(bg_rgba: np.ndarray, canvas_size: Tuple[int, int]) -> np.ndarray:
    """fill_gradient (background_resizing.py:63-98) after the PNG decode."""
    left, right, top, bottom = edge_strip_median_colors(bg_rgba)
    hv = sum((a - b) ** 2 for a, b in zip(left, right))
    vv = sum((a - b) ** 2 for a, b in zip(top, bottom))
    W, H = canvas_size
    out = np.empty((H, W, 4), np.uint8)
    horizontal = hv <= vv
    c1, c2 = (left, right) if horizontal else (top, bottom)
    a1 = np.array(c1, np.int32)
    a2 = np.array(c2, np.int32)
    lib().orc_fill_gradient(_p(out), W, H, int(horizontal), _p(a1), _p(a2))
    return out
