"""CPU-only tests: host-side mirror of the reference interface, the C-ABI library's exports,
and the host coefficient builder (the one part of the product that is not a kernel)."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle
from image_transformation_b200 import _native, synth
from image_transformation_b200.compositor import resolve_placements

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "b200comp.h")).read()
    declared = set(re.findall(r"\b(b200comp_[a-z0-9_]+)\s*\(", header))
    assert declared, "no prototypes parsed"
    assert declared == set(_native.EXPORTED)
    L = _native.lib()
    for name in sorted(declared):
        assert hasattr(L, name), name
    assert L.b200comp_abi_version() == 1
    assert L.b200comp_device_count() >= 0


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_native.Placement) == 48
    assert ctypes.sizeof(_native.Canvas) == 56


def test_no_gpu_fails_loudly():
    if _native.lib().b200comp_device_count() > 0:
        pytest.skip("a GPU is visible")
    from PIL import Image
    from image_transformation_b200.compositor import composite

    bg = Image.new("RGBA", (10, 10), (255, 0, 0, 255))
    obj = Image.new("RGBA", (2, 2), (0, 255, 0, 255))
    with pytest.raises(_native.B200CompError):
        composite(bg, {1: obj}, [{"object_id": 1, "box": [4, 4, 6, 6]}])


@pytest.mark.parametrize("pair", [(64, 32), (33, 64), (1536, 1), (100, 99), (7, 7), (256, 255), (1024, 513),
                                  (1365, 1024), (3, 400), (400, 3), (1, 1), (2, 1), (1, 2), (977, 611)])
def test_coefficient_builder_matches_oracle(pair):
    in_size, out_size = pair
    L = _native.lib()
    ks = L.b200comp_ksize(in_size, out_size)
    k = np.full((out_size, ks), 12345, np.int32)
    b = np.zeros((out_size, 2), np.int32)
    got = ctypes.c_int(0)
    assert L.b200comp_build_coeffs(in_size, out_size, k.ctypes.data, b.ctypes.data, ctypes.byref(got)) == 0
    ko, bo, kso = oracle.coeffs(in_size, out_size)
    assert got.value == ks == kso
    assert np.array_equal(k, ko) and np.array_equal(b, bo)


def test_coefficient_builder_rejects_bad_sizes():
    L = _native.lib()
    assert L.b200comp_ksize(0, 5) == -1
    assert "positive" in _native.last_error()
    assert L.b200comp_build_coeffs(4, 4, None, None, None) == -1


def test_resolve_placements_reference_semantics():
    sizes = {1: (10, 20), 2: (3, 400)}
    pl = [
        {"object_id": "1", "box": [1.9, -2.9, 11.2, 17.5]},   # str id, float box -> int() truncation toward 0
        {"object_id": 7, "box": "not even looked at"},        # unknown id skipped before the box is read
        {"object_id": 1, "box": [5, 5, 5, 2]},                # degenerate -> 1x1
        {"object_id": 2, "box": [0, 0, 9, 2]},                # tall image, shrinking height -> vertical first
        {"object_id": True, "box": [0, 0, 1, 1]},             # bool is an int: key 1
    ]
    r = resolve_placements(pl, sizes)
    assert r[0] == (1, 1, -2, 10, 19, 0)
    assert r[1] == (1, 5, 5, 1, 1, 0)
    assert r[2] == (2, 0, 0, 9, 2, _native.VERTICAL_FIRST)
    assert r[3][0] is True or r[3][0] == 1
    with pytest.raises(KeyError):
        resolve_placements([{"object_id": 1}], sizes)
    with pytest.raises(KeyError):
        resolve_placements([{"box": [0, 0, 1, 1]}], sizes)
    with pytest.raises(ValueError):
        resolve_placements([{"object_id": "a", "box": [0, 0, 1, 1]}], sizes)
    with pytest.raises(ValueError):
        resolve_placements([{"object_id": 1, "box": [0, 0, 1]}], sizes)
    with pytest.raises((ValueError, TypeError)):
        resolve_placements([{"object_id": 1, "box": [0, 0, 1, None]}], sizes)
    # Python ints do not wrap into the int32 fields of the C ABI: Pillow's own failures for such boxes
    # (Image.resize / alpha_composite argument parsing: OverflowError; an unallocatable target: MemoryError)
    with pytest.raises(OverflowError):
        resolve_placements([{"object_id": 1, "box": [0, 0, 2**31, 10]}], sizes)
    with pytest.raises(OverflowError):
        resolve_placements([{"object_id": 1, "box": [2**40, 0, 2**40 + 5, 10]}], sizes)
    with pytest.raises(OverflowError):
        resolve_placements([{"object_id": 1, "box": [-2**40, 0, -2**40 + 5, 10]}], sizes)
    with pytest.raises(MemoryError):
        resolve_placements([{"object_id": 1, "box": [0, 0, 2**30 + 5, 2]}], sizes)
    assert resolve_placements([{"object_id": 1, "box": [2**30 + 1, 0, 2**30 + 11, 10]}], sizes) == []  # clipped away entirely


def test_reference_placer_layout_fixture_is_in_spec():
    """tests/golden/c3_reference_layouts.npz (reference placer on size proxies): 20 boxes per canvas, inside the
    canvas, isotropic scales 0.5..1 of the pool cutouts -- and the same object draws as the in-repo generator."""
    pool = synth.workload_pool("c3_refplacer")
    sizes = {k: (v.shape[1], v.shape[0]) for k, v in pool.items()}
    for i in (0, 3, 255, 256 + 3):
        ref = synth.workload_placements("c3_refplacer", sizes, i)
        own = synth.workload_placements("c3_4k_20obj", sizes, i % 256)
        assert len(ref) == 20
        assert sorted(p["object_id"] for p in ref) == sorted(p["object_id"] for p in own)
        for p in ref:
            x1, y1, x2, y2 = p["box"]
            sw, sh = sizes[p["object_id"]]
            assert 0 <= x1 < x2 <= 3840 and 0 <= y1 < y2 <= 2160
            assert 0.5 - 0.01 <= (x2 - x1) / sw <= 1.0 + 0.01 and abs((x2 - x1) / sw - (y2 - y1) / sh) < 0.01


def test_synthetic_workload_is_deterministic_and_in_spec():
    sizes = {i: (300 + 7 * i, 900 - 5 * i) for i in range(1, 9)}
    a = synth.canvas_placements(sizes, (3840, 2160), 5)
    b = synth.canvas_placements(sizes, (3840, 2160), 5)
    assert a == b and len(a) == 20
    for p in a:
        x1, y1, x2, y2 = p["box"]
        sw, sh = sizes[p["object_id"]]
        assert 0.49 <= (x2 - x1) / sw <= 1.01 and x1 >= 0 and y1 >= 0
    u = synth.canvas_placements(sizes, (7680, 4320), 0, n_objects=64, layout="uniform")
    assert len(u) == 64
    st = synth.alpha_stats(synth.make_pool(4, 64, 128))
    assert 0.2 < st["transparent"] < 0.4 and 0.5 < st["opaque"] < 0.75


def test_unpremultiply_magic_division_is_exact():
    """kernels.cuh unpremultiply_px replaces 255 * c / a (Convert.c rgba2rgbA) by a multiply-high with
    m = ceil(2^24 / a): (255 * c << 8) * m >> 32 must be the truncated quotient for every (c, a)."""
    for a in range(1, 256):
        m = ((1 << 24) + a - 1) // a
        assert m <= 1 << 24
        for c in range(256):
            assert ((c * 65280) * m) >> 32 == (255 * c) // a, (c, a)


def test_rgba_array_paths():
    """_native.rgba_array / new_rgba_image / image_from_rgba: zero-copy views of single-block images stay coherent
    with the image, bigger and buffer-backed images are copied correctly, results are fully mutable PIL images."""
    from PIL import Image

    from image_transformation_b200 import _native

    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (60, 40, 4), dtype=np.uint8)
    for img in (Image.fromarray(a, "RGBA"), Image.fromarray(a, "RGBA").copy()):
        v = _native.rgba_array(img)
        assert v.shape == (60, 40, 4) and v.dtype == np.uint8 and np.array_equal(v, a)
    owned = Image.fromarray(a, "RGBA").copy()
    owned.putpixel((3, 5), (9, 8, 7, 6))  # a view (or a copy taken now) shows the pixels as they are at call time
    assert tuple(_native.rgba_array(owned)[5, 3]) == (9, 8, 7, 6)
    big = Image.new("RGBA", (3000, 1500), (1, 2, 3, 4))  # 18 MB: more than one 16 MB block
    big.putpixel((2999, 1499), (5, 6, 7, 8))
    v = _native.rgba_array(big)
    assert v.shape == (1500, 3000, 4) and tuple(v[1499, 2999]) == (5, 6, 7, 8) and tuple(v[0, 0]) == (1, 2, 3, 4)
    img = _native.image_from_rgba(a.copy())
    assert img.mode == "RGBA" and img.size == (40, 60) and np.array_equal(np.asarray(img), a)
    # a real mutable image: pixel-access writes, putpixel and paste all work in place (the reference's result is
    # `background_img.copy()`, compositor.py:11)
    assert not getattr(img, "readonly", 0)
    px = img.load()
    px[0, 0] = (1, 1, 1, 1)
    img.putpixel((1, 0), (2, 2, 2, 2))
    img.paste((3, 3, 3, 3), (2, 0, 3, 1))
    assert img.getpixel((0, 0)) == (1, 1, 1, 1) and img.getpixel((1, 0)) == (2, 2, 2, 2) and img.getpixel((2, 0)) == (3, 3, 3, 3)
    fresh, view = _native.new_rgba_image(7, 5)
    if fresh is not None:
        view[...] = 11
        assert fresh.getpixel((6, 4)) == (11, 11, 11, 11)
        fresh.putpixel((0, 0), (1, 2, 3, 4))
        assert tuple(view[0, 0]) == (1, 2, 3, 4)  # the view IS the image's memory
    # pixel_block / new_rgba_block: the raw address the drop-in hands to the C ABI (no numpy in between)
    import ctypes

    blk = _native.pixel_block(owned)
    if blk is not None:
        ptr, keep = blk
        assert ptr == _native.data_ptr(_native.rgba_array(owned)) and keep is not None
        first = (ctypes.c_uint8 * 4).from_address(ptr)
        assert tuple(first) == owned.getpixel((0, 0))
        owned.putpixel((0, 0), (4, 3, 2, 1))
        assert tuple(first) == (4, 3, 2, 1)
        img2, blk2 = _native.new_rgba_block(9, 4)
        ctypes.memset(blk2[0], 7, 9 * 4 * 4)
        assert img2.getpixel((8, 3)) == (7, 7, 7, 7) and not getattr(img2, "readonly", 0)
    assert _native.pixel_block(Image.new("RGB", (4, 4))) is None  # only RGBA blocks are handed over
    ro = Image.frombuffer("RGBA", (40, 60), a.tobytes(), "raw", "RGBA", 0, 1)  # buffer-backed, read-only: copied, not viewed
    assert _native.pixel_block(ro) is None and np.array_equal(_native.rgba_array(ro), a)
