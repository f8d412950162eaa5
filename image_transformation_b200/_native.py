"""ctypes binding of libb200comp.so (the C ABI declared in include/b200comp.h).

There is NO CPU fallback: importing works without a GPU (so host logic and symbol
checks can run anywhere), but every compute entry point raises if no CUDA device is
visible or the library is missing.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_int, c_int32, c_int64, c_size_t, c_uint32, c_void_p

import numpy as np
from PIL import Image

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
    "b200comp_composite_host", "b200comp_composite_host_ex", "b200comp_device_upload", "b200comp_device_free",
    "b200comp_trim", "b200comp_host_alloc", "b200comp_host_free", "b200comp_masked_median_rgb",
    "b200comp_fill_rgba", "b200comp_fill_gradient", "b200comp_masked_median_rgb_host",
    "b200comp_edge_strip_medians_host", "b200comp_fill_solid_host", "b200comp_fill_gradient_host",
)

VERTICAL_FIRST = 1
SRC_DEVICE = 2
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
    L.b200comp_composite_host_ex.argtypes = [vp, c_uint32, c_int, c_int, c_size_t, vp, c_size_t, POINTER(Placement), c_int]
    L.b200comp_device_upload.argtypes = [vp, c_int, c_int, c_size_t, POINTER(vp), POINTER(c_size_t)]
    L.b200comp_device_free.argtypes = [vp]
    L.b200comp_trim.argtypes = []
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


def device_count_quiet() -> int:
    """Visible CUDA devices, 0 if the library cannot be loaded (never raises)."""
    try:
        return int(_load().b200comp_device_count())
    except Exception:
        return 0


_gpu_seen = False


def require_gpu() -> None:
    global _gpu_seen
    if _gpu_seen:  # devices do not go away; the C entry points still fail loudly if one does
        return
    if _load().b200comp_device_count() < 1:
        raise B200CompError(
            "no CUDA device visible: the B200 compositor has no CPU fallback "
            "(use the reference's own compositor.py on machines without a GPU)"
        )
    _gpu_seen = True


# ---------------------------------------------------------------------------- PIL <-> memory
# Pillow >= 11.2 exports an image's own pixel memory through the Arrow C data interface (zero copy for images that live
# in one block).  The helpers below use it in both directions: inputs are read where they are (no tobytes(): that
# costs 40-70 ms for a 4K RGBA image), results are written by the library straight into a single-block image
# Pillow owns -- a real, fully mutable PIL image (pixel access writes, paste, ImageDraw all work in place).
class _ArrowArray(ctypes.Structure):
    pass


_ArrowArray._fields_ = [("length", c_int64), ("null_count", c_int64), ("offset", c_int64), ("n_buffers", c_int64),
                        ("n_children", c_int64), ("buffers", POINTER(c_void_p)),
                        ("children", POINTER(POINTER(_ArrowArray))), ("dictionary", c_void_p), ("release", c_void_p),
                        ("private_data", c_void_p)]
_capsule_ptr = ctypes.pythonapi.PyCapsule_GetPointer
_capsule_ptr.restype = c_void_p
_capsule_ptr.argtypes = [ctypes.py_object, c_char_p]


_memview = ctypes.pythonapi.PyMemoryView_FromMemory
_memview.restype = ctypes.py_object
_memview.argtypes = [c_void_p, ctypes.c_ssize_t, ctypes.c_int]
_PyBUF_WRITE = 0x200


def _make_view_type():
    class _PixelView(np.ndarray):
        # attributes: _b200_capsules (owner of the memory), ptr (address of pixel (0, 0))
        pass

    return _PixelView


_PixelViewType = None


def pixel_block(img):
    """(address, owner) of a single-block RGBA image's own pixel memory -- rows of width * 4 bytes, no padding -- or
    None if Pillow cannot export it without a copy (older Pillow, image spread over several blocks, buffer-backed
    image).  `owner` (an Arrow capsule) holds a reference to the block: keep it for as long as the address is used."""
    if img.mode != "RGBA" or getattr(img, "readonly", 0):
        return None
    try:
        img.load()
        capsule = img.im.__arrow_c_array__()  # (Image.__arrow_c_array__ would also build a schema capsule nobody reads)
    except (ValueError, NotImplementedError, AttributeError):
        return None
    arr = _ArrowArray.from_address(_capsule_ptr(capsule, b"arrow_array"))
    if arr.n_children != 1 or arr.offset != 0:
        return None
    child = arr.children[0].contents
    w, h = img.size
    if child.length != w * h * 4 or child.offset != 0 or child.n_buffers < 2:
        return None
    ptr = child.buffers[1]
    return (ptr, capsule) if ptr else None


def _arrow_view(img):
    """(H, W, 4) uint8 numpy view of pixel_block(img), or None."""
    global _PixelViewType
    blk = pixel_block(img)
    if blk is None:
        return None
    if _PixelViewType is None:
        _PixelViewType = _make_view_type()
    w, h = img.size
    view = np.frombuffer(_memview(blk[0], w * h * 4, _PyBUF_WRITE), np.uint8).reshape(h, w, 4).view(_PixelViewType)
    view._b200_capsules = blk[1]  # lives as long as the view
    view.ptr = blk[0]
    return view


def data_ptr(a) -> int:
    """Address of element 0 of an array handed to the C ABI (views of image blocks carry it; ndarray.ctypes is slow)."""
    p = getattr(a, "ptr", None)
    return p if p is not None else a.ctypes.data


def new_rgba_block(w: int, h: int):
    """A fresh single-block RGBA image (uninitialised pixels) and pixel_block() of it, or (None, None)."""
    if hasattr(Image.core, "new_block"):
        img = Image.Image()._new(Image.core.new_block("RGBA", (w, h)))
        blk = pixel_block(img)
        if blk is not None:
            return img, blk
    return None, None


def new_rgba_image(w: int, h: int):
    """A fresh single-block RGBA image (uninitialised pixels) and the writable numpy view of its memory."""
    if hasattr(Image.core, "new_block"):
        img = Image.Image()._new(Image.core.new_block("RGBA", (w, h)))
        view = _arrow_view(img)
        if view is not None:
            return img, view
    return None, None


def rgba_array(img):
    """(H, W, 4) uint8 array of a PIL RGBA image for the duration of one library call: a zero-copy view of the image's
    own memory when Pillow can export it (single-block images: everything up to 16 MB, and every image this package
    returned), else a copy -- through a single-block image and Pillow's C paste when possible (about three times
    faster than np.asarray / tobytes()), np.asarray as the last resort."""
    if img.mode == "RGBA":
        view = _arrow_view(img)  # (loads the image)
        if view is not None:
            return view
        img.load()
        if hasattr(Image.core, "new_block") and img.size[0] > 0 and img.size[1] > 0:
            try:
                blk = Image.core.new_block("RGBA", img.size)
                blk.paste(img.im, (0, 0) + img.size)
                view = _arrow_view(Image.Image()._new(blk))
                if view is not None:
                    return view
            except (ValueError, TypeError, NotImplementedError):
                pass
    return np.ascontiguousarray(np.asarray(img), dtype=np.uint8)


def image_from_rgba(out):
    """PIL RGBA image holding the pixels of a (H, W, 4) uint8 array (one copy into an image Pillow owns; callers on
    the hot path avoid even that by letting the library write into new_rgba_image())."""
    h, w = out.shape[:2]
    img, view = new_rgba_image(w, h) if w > 0 and h > 0 else (None, None)
    if img is None:
        return Image.fromarray(out, "RGBA").copy()
    view[...] = out
    return img
