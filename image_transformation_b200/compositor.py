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

# Pillow >= 12 resizes very tall images vertical-first (PIL Image.py:2431-2435); Pillow 11.3 (the reference's pin,
# requirements.txt:29) does not.  The drop-in follows the Pillow that is installed next to it -- the reference's own
# arithmetic in that environment -- unless B200COMP_PILLOW_COMPAT=11 / 12 says otherwise.
def _tall_image_vertical_first() -> bool:
    forced = os.environ.get("B200COMP_PILLOW_COMPAT")
    if forced in ("11", "12"):
        return forced == "12"
    try:
        import PIL

        return int(PIL.__version__.split(".")[0]) >= 12
    except Exception:  # pragma: no cover
        return True


TALL_IMAGE_VERTICAL_FIRST = _tall_image_vertical_first()


_INT32_MAX = 2**31 - 1
_GEOMETRY_MAX = 2**30  # b200comp_plan_create's range check


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
        # Python ints are unbounded, the C ABI's are int32 (and its geometry wants |x|, |y|, w, h <= 2^30).  Pillow's own
        # C entry points fail the same way the checks below do: sizes or offsets beyond a C int -> OverflowError from
        # the argument parser, a resize target it cannot allocate -> MemoryError.
        for v in (w, h, x1, y1):
            if v > _INT32_MAX:
                raise OverflowError("signed integer is greater than maximum")
            if v < -_INT32_MAX - 1:
                raise OverflowError("signed integer is less than minimum")
        if w > _GEOMETRY_MAX or h > _GEOMETRY_MAX:
            raise MemoryError(f"resize target {w}x{h} is too large")
        if abs(x1) > _GEOMETRY_MAX or abs(y1) > _GEOMETRY_MAX:
            continue  # cannot touch any canvas (w, h <= 2^30): alpha_composite would clip it away entirely
        sw, sh = sizes[oid]
        flags = _native.VERTICAL_FIRST if (TALL_IMAGE_VERTICAL_FIRST and sh > 100 * sw and h < sh) else 0
        out.append((oid, x1, y1, w, h, flags))
    return out


# ---- device-resident cutouts of a bundle (macro_placement_test.py:1493, 1679 reload the same results.json before
# every composite of the refine loop) ----------------------------------------------------------------------------
_memcmp = ctypes.CDLL(None).memcmp
_memcmp.restype = ctypes.c_int
_memcmp.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]


class _DeviceCutout:
    """A decoded cutout uploaded once; freed with the cache entry.  `host` is the decoded image's own memory: an image
    handed to composite() uses the device copy only if its pixels still equal it (memcmp: a caller may have drawn on
    the cutout it was given)."""

    def __init__(self, arr: np.ndarray):
        self.host = arr
        dev, pitch = ctypes.c_void_p(), ctypes.c_size_t()
        rc = _native.lib().b200comp_device_upload(arr.ctypes.data, arr.shape[1], arr.shape[0], arr.strides[0],
                                                  ctypes.byref(dev), ctypes.byref(pitch))
        _native.check(rc, "load_object_images")
        self.ptr, self.pitch, self.size = dev.value, pitch.value, (arr.shape[1], arr.shape[0])

    def __del__(self):
        try:
            if self.ptr:
                _native.lib().b200comp_device_free(self.ptr)
        except Exception:  # interpreter shutdown
            pass
        self.ptr = None


_CUTOUT_CACHE: Dict[Tuple[str, float, int], Dict[int, Tuple[Image.Image, "_DeviceCutout"]]] = {}
_CUTOUT_CACHE_MAX = 8  # bundles
CUTOUT_CACHE_STATS = {"uploads": 0, "hits": 0}


def invalidate_cutout_cache() -> None:
    """Forget every cached bundle (decoded cutouts and their device copies)."""
    _CUTOUT_CACHE.clear()


def _cutout_cache_enabled() -> bool:
    return os.environ.get("B200COMP_CUTOUT_CACHE", "1") != "0"


def _pixels(img: Image.Image):
    """(address, pitch, keepalive) of an RGBA image's pixels for one library call: the image's own block when Pillow
    exports it (no copy, no numpy), else a contiguous copy."""
    blk = _native.pixel_block(img)
    if blk is not None:
        return blk[0], img.size[0] * 4, blk[1]
    a = _native.rgba_array(img)
    return _native.data_ptr(a), a.strides[0], a


_u32_at = ctypes.c_uint32.from_address


def _solid_colour_of(img: Image.Image, ptr: int, nbytes: int):
    """RGBA value (little-endian uint32) of a canvas made by fill_solid() that is still all one colour, else None.
    `ptr` / `nbytes`: the canvas pixels, tightly packed.  The tag alone is not trusted: the caller may have drawn on
    the canvas since.  One memcmp of the buffer against itself shifted by a pixel proves every pixel equals the first
    (cheaper than the upload it saves)."""
    colour = getattr(img, "_b200_solid", None)
    if colour is None or nbytes < 4:
        return None
    if _u32_at(ptr).value != colour:
        return None
    if nbytes > 4 and _memcmp(ptr, ptr + 4, nbytes - 4) != 0:
        return None
    return colour


def composite(background_img: Image.Image, object_images: Dict[int, Image.Image], placements: List[Dict]) -> Image.Image:
    """Composite objects onto the background according to placements (compositor.py:6-22).

    placements: list of {object_id, box: [x1, y1, x2, y2]}; list order is z-order.
    The background is not modified; a new RGBA image is returned (a real, mutable PIL image).
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
    bg_ptr, bg_pitch, bg_keep = _pixels(background_img)
    result, blk = _native.new_rgba_block(W, H)  # the library writes straight into an image Pillow owns
    if result is not None:
        out, out_ptr, out_pitch = None, blk[0], W * 4
    else:  # Pillow without the Arrow export: plain array, one more copy at the end
        out = np.empty((H, W, 4), np.uint8)
        out_ptr, out_pitch = out.ctypes.data, out.strides[0]
    pixels: Dict[int, tuple] = {}
    keep = [bg_keep, blk]
    recs = (_native.Placement * len(resolved))()
    for i, (oid, x, y, w, h, flags) in enumerate(resolved):
        img = object_images[oid]
        px = pixels.get(oid)
        if px is None:
            px = pixels[oid] = _pixels(img)
        ptr, pitch, _ = px
        sw, sh = img.size
        dev = getattr(img, "_b200_dev", None)
        if dev is not None and dev.ptr and pitch == sw * 4 and dev.host.shape == (sh, sw, 4) and dev.host.flags.c_contiguous \
                and _memcmp(ptr, _native.data_ptr(dev.host), sw * sh * 4) == 0:
            # uploaded by load_object_images and still pixel for pixel what was uploaded
            keep.append(dev)
            recs[i] = _native.Placement(dev.ptr, dev.pitch, sw, sh, x, y, w, h, flags | _native.SRC_DEVICE, 0)
            continue
        recs[i] = _native.Placement(ptr, pitch, sw, sh, x, y, w, h, flags, 0)
    solid = _solid_colour_of(background_img, bg_ptr, W * H * 4) if bg_pitch == W * 4 else None
    if solid is not None:  # fill_solid() canvas, untouched: synthesised on the device from the colour
        rc = _native.lib().b200comp_composite_host_ex(None, solid, W, H, 0, out_ptr, out_pitch, recs, len(resolved))
    else:
        rc = _native.lib().b200comp_composite_host_ex(bg_ptr, 0, W, H, bg_pitch, out_ptr, out_pitch, recs, len(resolved))
    _native.check(rc, "composite")
    del keep, pixels
    return result if result is not None else _native.image_from_rgba(out)


def load_object_images(results_json_path: str) -> Dict[int, Image.Image]:
    """{object_id: RGBA cutout} from a bundle's results.json (compositor.py:25-35).  PNG decode stays on the host.

    With a GPU present the decoded bundle is cached by (path, mtime, size) together with a device copy of every
    cutout: the refine loop reloads the same bundle before each composite (macro_placement_test.py:1493, 1679), so
    from the second iteration on nothing is decoded or uploaded.  Every call returns fresh, fully mutable copies, as
    the reference does; composite() uses a cutout's device copy only while its pixels are still the uploaded ones.
    ``B200COMP_CUTOUT_CACHE=0`` or ``invalidate_cutout_cache()`` turn the cache off.
    """
    def decode() -> Dict[int, Image.Image]:
        with open(results_json_path, "r", encoding="utf-8") as f:
            items = json.load(f)
        root = os.path.dirname(results_json_path)
        return {int(it["object_id"]): Image.open(os.path.join(root, it["filename"])).convert("RGBA") for it in items}

    if not _cutout_cache_enabled() or _native.device_count_quiet() < 1:
        return decode()
    try:
        st = os.stat(results_json_path)
        root = os.path.dirname(results_json_path)
        with open(results_json_path, "r", encoding="utf-8") as f:
            names = [it["filename"] for it in json.load(f)]
        stamp = tuple((n, os.stat(os.path.join(root, n)).st_mtime_ns, os.stat(os.path.join(root, n)).st_size) for n in names)
    except (OSError, KeyError, TypeError, ValueError):
        return decode()  # let the reference's own errors surface from the plain path
    key = (os.path.abspath(results_json_path), st.st_mtime_ns, st.st_size, stamp)
    entry = _CUTOUT_CACHE.get(key)
    if entry is None:
        images = decode()
        entry = {}
        for oid, img in images.items():
            dev = None
            if img.size[0] > 0 and img.size[1] > 0:
                dev = _DeviceCutout(_native.rgba_array(img))
                CUTOUT_CACHE_STATS["uploads"] += 1
            entry[oid] = (img, dev)
        while len(_CUTOUT_CACHE) >= _CUTOUT_CACHE_MAX:
            _CUTOUT_CACHE.pop(next(iter(_CUTOUT_CACHE)))
        _CUTOUT_CACHE[key] = entry
    else:
        CUTOUT_CACHE_STATS["hits"] += 1
    out: Dict[int, Image.Image] = {}
    for oid, (img, dev) in entry.items():
        view = img.copy()  # a new image object per call, as the reference returns
        if dev is not None:
            view._b200_dev = dev
        out[oid] = view
    return out
