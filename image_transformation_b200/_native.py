"""ctypes binding of libb200comp.so (the C ABI declared in include/b200comp.h).

There is NO CPU fallback: importing works without a GPU (so host logic and symbol
checks can run anywhere), but every compute entry point raises if no CUDA device is
visible or the library is missing.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_int, c_int32, c_int64, c_size_t, c_uint32, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
# B200COMP_LIB selects another build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("B200COMP_LIB") or os.path.join(_PKG, "_lib", "libb200comp.so")

# every symbol include/b200comp.h declares (tests check the library exports all of them)
EXPORTED = (
    "b200comp_abi_version", "b200comp_last_error", "b200comp_device_count", "b200comp_ksize",
    "b200comp_build_coeffs", "b200comp_resize_rgba_lanczos", "b200comp_resample_rgba", "b200comp_alpha_over",
    "b200comp_plan_create", "b200comp_plan_run", "b200comp_plan_prepare", "b200comp_plan_run_canvases",
    "b200comp_plan_destroy", "b200comp_plan_info", "b200comp_plan_profile", "b200comp_plan_profile_read",
    "b200comp_plan_last_records",
    "b200comp_plan_check", "b200comp_composite_batch", "b200comp_composite_batch_host",
    "b200comp_composite_host", "b200comp_host_alloc", "b200comp_host_free", "b200comp_masked_median_rgb",
    "b200comp_fill_rgba", "b200comp_fill_gradient", "b200comp_masked_median_rgb_host",
    "b200comp_edge_strip_medians_host", "b200comp_fill_solid_host", "b200comp_fill_gradient_host",
)

VERTICAL_FIRST = 1
INFO_KEYS = ("algorithmic_bytes", "launches_per_run", "fused_placements", "identity_placements",
             "preresampled_placements", "coeff_bytes", "smem_bytes", "tiles")


class Placement(ctypes.Structure):
    """struct b200comp_placement"""
    _fields_ = [("src", c_void_p), ("src_pitch", c_int64), ("sw", c_int32), ("sh", c_int32), ("x", c_int32),
                ("y", c_int32), ("w", c_int32), ("h", c_int32), ("flags", c_int32), ("reserved", c_int32)]


class Canvas(ctypes.Structure):
    """struct b200comp_canvas"""
    _fields_ = [("out", c_void_p), ("out_pitch", c_int64), ("bg", c_void_p), ("bg_pitch", c_int64),
                ("solid_rgba", c_uint32), ("W", c_int32), ("H", c_int32), ("first_placement", c_int32),
                ("n_placements", c_int32), ("reserved", c_int32)]


class B200CompError(RuntimeError):
    pass


_lib = None


def _load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        # build in-tree if a toolchain is present; otherwise fail loudly (no fallback path exists)
        from . import build as _build

        try:
            _build.build()
        except Exception as exc:  # pragma: no cover - depends on the toolchain
            raise ImportError(
                f"libb200comp.so is missing ({LIB_PATH}) and could not be built: {exc}. "
                "Run `python -m image_transformation_b200.build`; there is no CPU fallback."
            ) from exc
    L = ctypes.CDLL(LIB_PATH)
    vp, i32p = c_void_p, POINTER(c_int32)
    L.b200comp_abi_version.restype = c_int
    L.b200comp_last_error.restype = c_char_p
    L.b200comp_device_count.restype = c_int
    L.b200comp_ksize.argtypes = [c_int, c_int]
    L.b200comp_build_coeffs.argtypes = [c_int, c_int, vp, vp, POINTER(c_int)]
    L.b200comp_resize_rgba_lanczos.argtypes = [vp, c_int, c_int, c_size_t, vp, c_int, c_int, c_size_t, c_int, vp]
    L.b200comp_resample_rgba.argtypes = [vp, c_int, c_int, c_size_t, vp, c_int, c_int, c_size_t, vp, vp, c_int, vp,
                                         vp, c_int, vp, c_int, vp]
    L.b200comp_alpha_over.argtypes = [vp, c_int, c_int, c_size_t, vp, c_int, c_int, c_size_t, c_int, c_int, vp]
    L.b200comp_plan_create.argtypes = [POINTER(Canvas), c_int, POINTER(Placement), c_int, c_int, vp, POINTER(vp)]
    L.b200comp_plan_run.argtypes = [vp, vp]
    L.b200comp_plan_prepare.argtypes = [vp, vp]
    L.b200comp_plan_run_canvases.argtypes = [vp, c_int, c_int, vp]
    L.b200comp_plan_destroy.argtypes = [vp]
    L.b200comp_plan_info.argtypes = [vp, POINTER(c_int64)]
    L.b200comp_plan_check.argtypes = [vp, vp]
    L.b200comp_plan_last_records.argtypes = [vp, vp, POINTER(c_int64)]
    L.b200comp_plan_profile.argtypes = [vp, c_int]
    L.b200comp_plan_profile_read.argtypes = [vp, POINTER(ctypes.c_double), POINTER(c_int)]
    L.b200comp_composite_batch.argtypes = [POINTER(Canvas), c_int, POINTER(Placement), c_int, vp]
    L.b200comp_composite_batch_host.argtypes = [POINTER(Canvas), c_int, POINTER(Placement), c_int, c_int, c_int, c_int]
    L.b200comp_composite_host.argtypes = [vp, c_int, c_int, c_size_t, vp, c_size_t, POINTER(Placement), c_int]
    L.b200comp_host_alloc.argtypes = [POINTER(vp), c_size_t]
    L.b200comp_host_free.argtypes = [vp]
    L.b200comp_masked_median_rgb.argtypes = [vp, c_int, c_int, c_size_t, c_int, c_int, c_int, c_int, i32p, vp]
    L.b200comp_fill_rgba.argtypes = [vp, c_int, c_int, c_size_t, c_uint32, vp]
    L.b200comp_fill_gradient.argtypes = [vp, c_int, c_int, c_size_t, c_int, i32p, i32p, vp]
    L.b200comp_masked_median_rgb_host.argtypes = [vp, c_int, c_int, c_size_t, c_int, c_int, c_int, c_int, i32p]
    L.b200comp_edge_strip_medians_host.argtypes = [vp, c_int, c_int, c_size_t, c_int, i32p]
    L.b200comp_fill_solid_host.argtypes = [vp, c_int, c_int, c_size_t, vp, c_int, c_int, c_size_t, i32p]
    L.b200comp_fill_gradient_host.argtypes = [vp, c_int, c_int, c_size_t, vp, c_int, c_int, c_size_t, c_int, i32p,
                                              POINTER(c_int)]
    for name in EXPORTED:
        fn = getattr(L, name)
        if fn.restype is c_int and name not in ("b200comp_abi_version", "b200comp_device_count"):
            fn.restype = c_int
    _lib = L
    return L


def lib() -> ctypes.CDLL:
    return _load()


def last_error() -> str:
    msg = _load().b200comp_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, what: str = "b200comp") -> None:
    """Map a negative status to the Python exception the reference's callers would see."""
    if rc >= 0:
        return
    msg = f"{what}: {last_error()}"
    if rc == -1:
        raise ValueError(msg)
    if rc == -3:
        raise MemoryError(msg)
    raise B200CompError(msg)


def require_gpu() -> None:
    if _load().b200comp_device_count() < 1:
        raise B200CompError(
            "no CUDA device visible: the B200 compositor has no CPU fallback "
            "(use the reference's own compositor.py on machines without a GPU)"
        )


def rgba_array(img):
    """(H, W, 4) uint8 array of a PIL RGBA image: the array the image was built on when this package produced it
    (image_from_rgba) and nobody wrote to it since, else a copy (np.asarray, i.e. Pillow's tobytes()).
    Pillow >= 11.2 can also export single-block images zero-copy through Arrow; that path was measured (0.2 ms
    instead of ~1 ms per MB) but is not used: Pillow 12.2's export crashes the process on buffer-backed images.
    The result is only read during the call that asked for it."""
    import numpy as np

    w, h = img.size
    tagged = getattr(img, "_b200_rgba", None)
    if tagged is not None and getattr(img, "readonly", 0) and tagged.shape == (h, w, 4):
        return tagged
    return np.ascontiguousarray(np.asarray(img), dtype=np.uint8)


def image_from_rgba(out):
    """PIL RGBA image over a freshly produced (H, W, 4) uint8 array without a second copy of the pixels
    (Image.fromarray would copy them again: 33 MB at 4K).  The image is a real mutable PIL image: Pillow marks
    buffer-backed images read-only and copies on the first write (putpixel, paste, alpha_composite, ImageDraw)."""
    from PIL import Image

    h, w = out.shape[:2]
    img = Image.frombuffer("RGBA", (w, h), out, "raw", "RGBA", 0, 1)
    img._b200_rgba = out  # lets a later composite() / statistics call on this image skip the PIL -> NumPy copy
    return img
