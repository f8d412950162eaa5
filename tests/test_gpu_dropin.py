"""GPU tests of the drop-in boundary (PIL in / PIL out) beyond the golden cases of test_gpu_parity.py: the
reference's own test file run unmodified against the drop-in modules, mutability of the returned images, the bundle
loader and its device cutout cache around the refine loop (macro_placement_test.py:1493-1513, 1679-1699), concurrent
callers, the Pillow-11 resize order, and larger full-size samples of BASELINE.json configs[2..4]."""
import json
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

import golden_io as G
import oracle

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def assert_same(got, exp, what):
    got, exp = np.asarray(got), np.asarray(exp)
    assert got.shape == exp.shape, f"{what}: shape {got.shape} != {exp.shape}"
    if not np.array_equal(got, exp):
        bad = np.argwhere((got != exp).any(axis=-1))
        raise AssertionError(f"{what}: {len(bad)} pixels differ, first at {bad[0].tolist()}: "
                             f"{got[tuple(bad[0])].tolist()} != {exp[tuple(bad[0])].tolist()}")


def pil(a):
    from PIL import Image

    return Image.fromarray(np.ascontiguousarray(a)).copy()


def write_bundle(tmp_path, cutouts):
    """A bundle directory like output/squarespace: <name>.png per cutout + results.json (compositor.py:25-35)."""
    items = []
    for oid, a in cutouts.items():
        name = f"obj_{oid}.png"
        pil(a).save(os.path.join(tmp_path, name))
        items.append({"object_id": oid, "filename": name, "label": f"thing {oid}"})
    path = os.path.join(tmp_path, "results.json")
    with open(path, "w", encoding="utf-8") as f:
        json.dump(items, f)
    return path


# ------------------------------------------------------------------ the reference's own callers
@pytest.mark.skipif(not os.path.isdir(REF), reason="baseline/_ref missing (run baseline/install_ref.sh where /root/reference exists)")
def test_reference_test_file_runs_unmodified_against_the_dropin(tmp_path):
    """/root/reference/tests/test_compositor.py:1-11 imports `compositor` by name; with dropin/ first on sys.path it
    exercises this package.  The file is run as it is, from the unmodified copy in baseline/_ref."""
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "dropin"), ROOT, REF])
    probe = subprocess.run([sys.executable, "-c", "import compositor, background_resizing; print(compositor.__file__); "
                                                  "print(background_resizing.__file__)"],
                           env=env, capture_output=True, text=True, cwd=str(tmp_path), timeout=300)  # (cwd: '' heads sys.path)
    assert probe.returncode == 0, probe.stderr
    assert all(os.path.join(ROOT, "dropin") in line for line in probe.stdout.split()), probe.stdout
    run = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "--rootdir", str(tmp_path),
                          os.path.join(REF, "tests", "test_compositor.py")],
                         env=env, capture_output=True, text=True, cwd=str(tmp_path), timeout=600)
    assert run.returncode == 0 and "1 passed" in run.stdout, run.stdout + run.stderr


@pytest.mark.skipif(not os.path.isdir(REF), reason="baseline/_ref missing")
def test_reference_compose_block_through_the_dropin(tmp_path):
    """The compose block of run_macro_only (macro_placement_test.py:1427-1430, 1493-1513): fill_solid -> canvas.png ->
    load_object_images -> reference Flex placer -> composite -> save, with the drop-in modules; against the same block
    run on the reference's own modules (Pillow) in a subprocess."""
    script = r'''
import json, sys
from PIL import Image
from compositor import composite, load_object_images
from background_resizing import fill_solid
from layout_constraints import compute_canvas_size
bundle, out = sys.argv[1], sys.argv[2]
size = compute_canvas_size((970, 250), "1:1")
canvas = fill_solid(bundle + "/background.png", size)
canvas.save(out + "/canvas.png")
objects = load_object_images(bundle + "/results.json")
placements = [
    {"object_id": k, "box": [10 + 37 * i, 20 + 41 * i, 10 + 37 * i + int(im.size[0] * 0.8), 20 + 41 * i + int(im.size[1] * 0.8)]}
    for i, (k, im) in enumerate(sorted(objects.items()))]
bg = Image.open(out + "/canvas.png").convert("RGBA")
draft = composite(bg, objects, placements)
draft.save(out + "/draft.png")
'''
    bundle = os.path.join(REF, "output", "squarespace")
    outs = {}
    for name, path in (("ref", [REF]), ("dropin", [os.path.join(ROOT, "dropin"), ROOT, REF])):
        d = tmp_path / name
        d.mkdir()
        env = dict(os.environ)
        env["PYTHONPATH"] = os.pathsep.join(path)
        r = subprocess.run([sys.executable, "-c", script, bundle, str(d)], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        from PIL import Image

        outs[name] = (np.array(Image.open(d / "canvas.png")), np.array(Image.open(d / "draft.png")))
    assert_same(outs["dropin"][0], outs["ref"][0], "fill_solid canvas")
    assert_same(outs["dropin"][1], outs["ref"][1], "composited draft")


# ------------------------------------------------------------------ returned images are real PIL images
def test_results_are_fully_mutable(tmp_path):
    from PIL import Image, ImageDraw

    from image_transformation_b200.background_resizing import fill_gradient, fill_solid
    from image_transformation_b200.compositor import composite

    bg = Image.new("RGBA", (64, 48), (255, 0, 0, 255))
    obj = Image.new("RGBA", (8, 8), (0, 255, 0, 128))
    bgp = tmp_path / "background.png"
    Image.new("RGBA", (40, 30), (10, 20, 30, 255)).save(bgp)
    for img in (composite(bg, {1: obj}, [{"object_id": 1, "box": [4, 4, 20, 20]}]),
                fill_solid(str(bgp), (50, 20)), fill_gradient(str(bgp), (50, 20))):
        assert isinstance(img, Image.Image) and img.mode == "RGBA" and not getattr(img, "readonly", 0)
        before = np.array(img)
        px = img.load()
        px[0, 0] = (1, 2, 3, 4)  # raised "image is readonly" in round 1
        img.putpixel((1, 0), (5, 6, 7, 8))
        ImageDraw.Draw(img).point((2, 0), fill=(9, 9, 9, 9))
        img.alpha_composite(Image.new("RGBA", (2, 2), (0, 0, 255, 255)), dest=(5, 5))
        after = np.array(img)
        assert tuple(after[0, 0]) == (1, 2, 3, 4) and tuple(after[0, 1]) == (5, 6, 7, 8) and tuple(after[0, 2]) == (9, 9, 9, 9)
        assert tuple(after[5, 5]) == (0, 0, 255, 255)
        changed = (before != after).any(axis=-1)
        assert changed.sum() == 7  # exactly the pixels written
        # feeding a result back in composites onto what the image shows NOW
        out2 = composite(img, {1: obj}, [{"object_id": 1, "box": [0, 0, 8, 8]}])
        exp = oracle.composite(after, {1: np.array(obj)}, [{"object_id": 1, "box": [0, 0, 8, 8]}])
        assert_same(np.array(out2), exp, "composite onto a mutated result")
        c = img.copy()
        c.putpixel((3, 3), (1, 1, 1, 1))
        assert img.getpixel((3, 3)) != (1, 1, 1, 1)


def test_fill_solid_canvas_is_passed_by_colour_until_it_is_drawn_on(tmp_path):
    """fill_solid -> composite (macro_placement_test.py:1497-1511): the canvas colour travels as a value, the W*H*4
    bytes are not uploaded -- but only while the canvas still is that colour."""
    from PIL import Image

    from image_transformation_b200 import _native
    from image_transformation_b200 import compositor as C
    from image_transformation_b200.background_resizing import fill_solid

    rng = np.random.default_rng(5)
    bgp = tmp_path / "background.png"
    Image.fromarray(rng.integers(0, 256, (30, 40, 4), dtype=np.uint8)).save(bgp)
    obj = rng.integers(0, 256, (90, 120, 4), dtype=np.uint8)
    pl = [{"object_id": 1, "box": [10, 5, 170, 125]}, {"object_id": 1, "box": [150, 60, 270, 150]}]
    canvas = fill_solid(str(bgp), (300, 200))
    def tagged(img, pixels_of):
        a = _native.rgba_array(pixels_of)
        return C._solid_colour_of(img, _native.data_ptr(a), a.nbytes)

    colour = tagged(canvas, canvas)
    assert colour is not None and colour >> 24 == 255
    exp = oracle.composite(np.array(canvas), {1: obj}, pl)
    assert_same(np.array(C.composite(canvas, {1: pil(obj)}, pl)), exp, "solid canvas by value")
    canvas.putpixel((299, 199), (1, 2, 3, 4))  # last pixel: the cheapest place to miss
    assert tagged(canvas, canvas) is None
    exp = oracle.composite(np.array(canvas), {1: obj}, pl)
    assert_same(np.array(C.composite(canvas, {1: pil(obj)}, pl)), exp, "drawn-on canvas by buffer")
    assert tagged(canvas.copy(), canvas) is None  # copies carry no tag


def test_load_object_images_and_cutout_cache_around_the_refine_loop(tmp_path):
    """compositor.py:25-35 through the drop-in: ids, modes, pixels; and the loop of macro_placement_test.py:1679-1699
    (reload the bundle, new layout, composite) uploads every cutout once."""
    from PIL import Image

    from image_transformation_b200 import compositor as C

    rng = np.random.default_rng(5)
    cut = {3: rng.integers(0, 256, (40, 60, 4), dtype=np.uint8), 11: rng.integers(0, 256, (25, 30, 4), dtype=np.uint8)}
    cut[11][..., 3] = 255
    path = write_bundle(str(tmp_path), cut)
    Image.fromarray(cut[11][..., :3]).save(tmp_path / "obj_11.png")  # an RGB png: the loader converts it to RGBA
    C.invalidate_cutout_cache()
    C.CUTOUT_CACHE_STATS.update(uploads=0, hits=0)
    bg = np.full((120, 160, 4), (220, 238, 245, 255), np.uint8)
    for it in range(5):
        objs = C.load_object_images(path)
        assert sorted(objs) == [3, 11] and all(isinstance(k, int) for k in objs)
        assert all(im.mode == "RGBA" for im in objs.values())
        assert_same(np.array(objs[3]), cut[3], "decoded cutout 3")
        assert_same(np.array(objs[11]), cut[11], "decoded cutout 11 (RGB file)")
        pl = [{"object_id": "3", "box": [5 + it, 7, 5 + it + 45, 7 + 30]}, {"object_id": 11, "box": [70.9, 20 + 2 * it, 130.2, 70]},
              {"object_id": 99, "box": [0, 0, 5, 5]}]
        out = C.composite(pil(bg), objs, pl)
        assert_same(np.array(out), oracle.composite(bg, cut, pl), f"iteration {it}")
    assert C.CUTOUT_CACHE_STATS == {"uploads": 2, "hits": 4}
    # a caller that draws on a cutout it was handed gets what it drew, not the cached upload
    objs = C.load_object_images(path)
    objs[3].putpixel((0, 0), (1, 2, 3, 255))
    mod = dict(cut)
    mod[3] = np.array(objs[3])
    pl = [{"object_id": 3, "box": [0, 0, 60, 40]}]
    assert_same(np.array(C.composite(pil(bg), objs, pl)), oracle.composite(bg, mod, pl), "mutated cutout")
    # a rewritten bundle is decoded and uploaded again
    cut2 = {3: rng.integers(0, 256, (40, 60, 4), dtype=np.uint8)}
    path = write_bundle(str(tmp_path), cut2)
    objs = C.load_object_images(path)
    assert sorted(objs) == [3]
    assert_same(np.array(objs[3]), cut2[3], "bundle rewritten")
    # the cache off: same results
    os.environ["B200COMP_CUTOUT_CACHE"] = "0"
    try:
        objs = C.load_object_images(path)
        assert not hasattr(objs[3], "_b200_dev")
        assert_same(np.array(C.composite(pil(bg), objs, pl)), oracle.composite(bg, cut2, pl), "cache off")
    finally:
        del os.environ["B200COMP_CUTOUT_CACHE"]
    with pytest.raises(FileNotFoundError):
        C.load_object_images(str(tmp_path / "missing.json"))


def test_concurrent_callers():
    """Streamlit runs one script thread per session (SURVEY 8b threading): four threads composite different golden
    cases at once, each on its own per-thread context, all bit-exact."""
    from image_transformation_b200.compositor import composite

    names = G.case_names()[:8]
    cases = [G.case(n) for n in names]
    errors = []

    def work(tid):
        try:
            for rep in range(6):
                k = (tid + rep) % len(cases)
                bg, objs, pl, exp = cases[k]
                out = composite(pil(bg), {i: pil(v) for i, v in objs.items()}, pl)
                assert_same(np.array(out), exp, f"thread {tid} case {names[k]}")
        except Exception as exc:  # noqa: BLE001
            errors.append(repr(exc))

    threads = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    assert not errors, errors


def test_4k_canvas_through_the_dropin_from_two_threads():
    """One full-size C3 canvas (3840x2160, 20 objects of 256..1536 px) through PIL in / PIL out: the host copies of
    such a call are spread over the library's helper threads (host_api.cu CopyPool); two callers at once share them."""
    from image_transformation_b200 import synth
    from image_transformation_b200.compositor import composite

    pool = synth.make_pool(12, 256, 1536, seed=1234)
    sizes_by_id = {k: (v.shape[1], v.shape[0]) for k, v in pool.items()}
    pls = [synth.canvas_placements(sizes_by_id, (3840, 2160), i) for i in (1, 2)]
    bgs = []
    for colour in ((220, 238, 245, 255), (12, 40, 90, 255)):
        b = np.empty((2160, 3840, 4), np.uint8)
        b[...] = colour
        b[::7, ::5, :3] ^= 0x55  # not a solid colour
        bgs.append(b)
    exp = [oracle.composite(b, pool, pl) for b, pl in zip(bgs, pls)]
    objs = {k: pil(v) for k, v in pool.items()}
    errors = []

    def work(t):
        try:
            for _ in range(3):
                assert_same(np.array(composite(pil(bgs[t]), objs, pls[t])), exp[t], f"4K canvas, thread {t}")
        except Exception as exc:  # noqa: BLE001
            errors.append(repr(exc))

    threads = [threading.Thread(target=work, args=(t,)) for t in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    assert not errors, errors


def test_solid_canvas_and_device_sources_through_the_c_abi():
    """b200comp_composite_host_ex: bg == NULL composites onto the solid colour (what fill_solid returned), and
    B200COMP_SRC_DEVICE placements read cutouts uploaded once with b200comp_device_upload."""
    import ctypes

    from image_transformation_b200 import _native

    L = _native.lib()
    rng = np.random.default_rng(9)
    W, H = 300, 200
    cut = {1: rng.integers(0, 256, (70, 90, 4), dtype=np.uint8), 2: rng.integers(0, 256, (50, 50, 4), dtype=np.uint8)}
    pl = [{"object_id": 1, "box": [10, 10, 130, 110]}, {"object_id": 2, "box": [100, 60, 150, 110]}, {"object_id": 1, "box": [-20, 150, 70, 220]}]
    solid = (38, 73, 115, 255)
    bg = np.full((H, W, 4), solid, np.uint8)
    exp = oracle.composite(bg, cut, pl)
    dev = {}
    for k, a in cut.items():
        p, pitch = ctypes.c_void_p(), ctypes.c_size_t()
        _native.check(L.b200comp_device_upload(a.ctypes.data, a.shape[1], a.shape[0], a.strides[0], ctypes.byref(p), ctypes.byref(pitch)))
        dev[k] = (p.value, pitch.value)
    try:
        for use_dev in (False, True):
            recs = (_native.Placement * len(pl))()
            for i, p in enumerate(pl):
                x1, y1, x2, y2 = p["box"]
                a = cut[p["object_id"]]
                if use_dev and p["object_id"] == 1:
                    recs[i] = _native.Placement(dev[1][0], dev[1][1], a.shape[1], a.shape[0], x1, y1, x2 - x1, y2 - y1, _native.SRC_DEVICE, 0)
                else:
                    recs[i] = _native.Placement(a.ctypes.data, a.strides[0], a.shape[1], a.shape[0], x1, y1, x2 - x1, y2 - y1, 0, 0)
            out = np.zeros((H, W, 4), np.uint8)
            rgba = solid[0] | (solid[1] << 8) | (solid[2] << 16) | (solid[3] << 24)
            _native.check(L.b200comp_composite_host_ex(None, rgba, W, H, 0, out.ctypes.data, out.strides[0], recs, len(pl)))
            assert_same(out, exp, f"solid canvas, device sources {use_dev}")
    finally:
        for p, _ in dev.values():
            L.b200comp_device_free(p)
    assert L.b200comp_trim() == 0
    for i, p in enumerate(pl):  # host sources again: the device copies are gone
        x1, y1, x2, y2 = p["box"]
        a = cut[p["object_id"]]
        recs[i] = _native.Placement(a.ctypes.data, a.strides[0], a.shape[1], a.shape[0], x1, y1, x2 - x1, y2 - y1, 0, 0)
    out = np.zeros((H, W, 4), np.uint8)  # the per-thread context comes back after a trim
    _native.check(L.b200comp_composite_host(bg.ctypes.data, W, H, bg.strides[0], out.ctypes.data, out.strides[0], recs, len(pl)))
    assert_same(out, exp, "after trim")


# ------------------------------------------------------------------ Pillow 11 / 12 resize order
def test_tall_cutout_in_both_pillow_modes(monkeypatch):
    """Pillow >= 12 resizes src_h > 100 * src_w downscales vertical-first (PIL Image.py:2431-2435); 11.3 (the
    reference's pin, requirements.txt:29) does not.  The drop-in follows the installed Pillow unless
    B200COMP_PILLOW_COMPAT says otherwise; both orders are checked against the oracle."""
    import PIL

    from image_transformation_b200 import compositor as C

    assert C.TALL_IMAGE_VERTICAL_FIRST == (int(PIL.__version__.split(".")[0]) >= 12)
    rng = np.random.default_rng(3)
    tall = rng.integers(0, 256, (400, 3, 4), dtype=np.uint8)
    bg = np.full((40, 60, 4), (200, 100, 50, 255), np.uint8)
    pl = [{"object_id": 1, "box": [10, 5, 19, 25]}]  # 3x400 -> 9x20
    outs = {}
    for vf in (True, False):
        monkeypatch.setattr(C, "TALL_IMAGE_VERTICAL_FIRST", vf)
        got = np.array(C.composite(pil(bg), {1: pil(tall)}, pl))
        exp = oracle.composite(bg, {1: tall}, pl, vertical_first_rule=vf)
        assert_same(got, exp, f"vertical_first_rule={vf}")
        outs[vf] = got
    assert not np.array_equal(outs[True], outs[False])  # the two Pillow generations really differ here
    if C.TALL_IMAGE_VERTICAL_FIRST:  # and the installed Pillow agrees with its mode
        from PIL import Image

        ref = pil(bg)
        ref.alpha_composite(pil(tall).resize((9, 20), Image.LANCZOS), dest=(10, 5))
        assert_same(outs[True], np.array(ref), "installed Pillow")


# ------------------------------------------------------------------ larger full-size samples
def test_c4_sixteen_canvases_per_aspect_ratio():
    """BASELINE.json configs[3]: 64 canvases (16 of each aspect ratio, the 4399-wide ones on the plain load / store
    path), bit-exact against the oracle."""
    import torch

    from image_transformation_b200 import batch as Bm
    from image_transformation_b200 import synth

    pool = synth.make_pool(12, 256, 1024, seed=77)
    sizes_by_id = {k: (v.shape[1], v.shape[0]) for k, v in pool.items()}
    base = synth.WORKLOADS["c4_aspect_sweep"]["canvases"]
    sizes = [base[i % 4] for i in range(64)]
    pls = [synth.canvas_placements(sizes_by_id, s, 1000 + i, n_objects=12) for i, s in enumerate(sizes)]
    dpool = Bm.CutoutPool(pool)
    cb = Bm.CompositeBatch(dpool, sizes, pls, solid=(38, 73, 115, 255))
    cb.run()
    cb.check()
    for i, ((W, H), pl) in enumerate(zip(sizes, pls)):
        bg = np.empty((H, W, 4), np.uint8)
        bg[...] = (38, 73, 115, 255)
        assert_same(cb.output(i).cpu().numpy(), oracle.composite(bg, pool, pl), f"C4 canvas {i} {W}x{H}")
    cb.close()
    torch.cuda.empty_cache()


def test_c5_three_8k_canvases_with_8k_background_statistics():
    """BASELINE.json configs[4]: the masked median of a 7680x4320 background (binary alpha, 30 % transparent) and
    three 8K canvases of 64 large overlapping objects, bit-exact against the oracle."""
    import torch

    from image_transformation_b200 import batch as Bm
    from image_transformation_b200 import synth

    bgsrc = synth.synthetic_background(7680, 4320)
    colour = Bm.masked_median_rgb(torch.from_numpy(bgsrc).cuda())
    assert list(colour) == list(oracle.masked_median_rgb(bgsrc))
    pool = synth.make_pool(6, 1024, 2048, seed=4322)
    sizes_by_id = {k: (v.shape[1], v.shape[0]) for k, v in pool.items()}
    W, H = synth.WORKLOADS["c5_8k_64obj"]["canvas"]
    pls = [synth.canvas_placements(sizes_by_id, (W, H), 50 + i, n_objects=64, scale_lo=0.6, scale_hi=1.0, layout="uniform") for i in range(3)]
    dpool = Bm.CutoutPool(pool)
    cb = Bm.CompositeBatch(dpool, [(W, H)] * 3, pls, solid=(*colour, 255))
    cb.run()
    cb.check()
    bg = np.empty((H, W, 4), np.uint8)
    bg[...] = (*colour, 255)
    for i in range(3):
        assert_same(cb.output(i).cpu().numpy(), oracle.composite(bg, pool, pls[i]), f"C5 canvas {i}")
    cb.close()
    torch.cuda.empty_cache()
