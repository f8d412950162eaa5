"""Pins the CPU oracle (oracle/compositor_oracle.c) against the golden vectors that
tests/golden/make_golden.py recorded from the unmodified reference (compositor.py,
background_resizing.py on Pillow 12.2.0 / NumPy 2.3.5).  CPU only."""
import numpy as np
import pytest

import golden_io as G
import oracle


def test_reference_known_answer():
    # /root/reference/tests/test_compositor.py:5-11
    bg, objs, pl, exp = G.case("reference_known_answer")
    out = oracle.composite(bg, objs, pl)
    assert tuple(out[4, 4, :3]) == (0, 255, 0)
    assert np.array_equal(out, exp)


@pytest.mark.parametrize("entry", G.manifest()["stage"]["resize"], ids=lambda e: e["key"])
def test_resize_matches_pillow(entry):
    src = G.stage(entry["key"] + "/src")
    exp = G.stage(entry["key"] + "/out")
    assert G.sha(exp) == entry["sha256"]
    out = oracle.resize_rgba_lanczos(src, tuple(entry["dst"]))
    assert np.array_equal(out, exp), f"max diff {np.abs(out.astype(int) - exp.astype(int)).max()}"


def test_tall_image_rule_matters():
    # Pillow 12.x vertical-first branch (PIL Image.py:2431-2435): the golden was made with it
    e = next(e for e in G.manifest()["stage"]["resize"] if e["src"] == [3, 400] and e["dst"] == [9, 2] and e["mode"] == "random")
    src = G.stage(e["key"] + "/src")
    exp = G.stage(e["key"] + "/out")
    assert np.array_equal(oracle.resize_rgba_lanczos(src, (9, 2), vertical_first_rule=True), exp)
    assert not np.array_equal(oracle.resize_rgba_lanczos(src, (9, 2), vertical_first_rule=False), exp)


def test_premultiply_exhaustive():
    assert np.array_equal(oracle.premultiply(G.stage("premul/in")), G.stage("premul/out"))


def test_unpremultiply_exhaustive():
    assert np.array_equal(oracle.unpremultiply(G.stage("unpremul/in")), G.stage("unpremul/out"))


@pytest.mark.parametrize("entry", G.manifest()["stage"]["over"], ids=lambda e: e["key"])
def test_alpha_over(entry):
    key = entry["key"]
    exp = G.stage(key + "/out")
    assert G.sha(exp) == entry["sha256"]
    if "grid" in key:
        sa, da = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8))
        s = np.zeros((256, 256, 4), np.uint8)
        d = np.zeros((256, 256, 4), np.uint8)
        s[..., :3] = entry["src_rgb"]
        d[..., :3] = entry["dst_rgb"]
        s[..., 3] = sa
        d[..., 3] = da
    else:
        s, d = G.stage(key + "/src"), G.stage(key + "/dst")
    canvas = d.copy()
    oracle.alpha_over_inplace(canvas, s, (0, 0))
    assert np.array_equal(canvas, exp)


@pytest.mark.parametrize("name", G.case_names())
def test_composite_cases(name):
    bg, objs, pl, exp = G.case(name)
    bg0 = bg.copy()
    out = oracle.composite(bg, objs, pl)
    assert np.array_equal(out, exp), f"{(out != exp).any(axis=-1).sum()} pixels differ"
    assert np.array_equal(bg, bg0)  # input not mutated (compositor.py:11)


@pytest.mark.parametrize("entry", G.manifest()["stage"]["median"], ids=lambda e: e["key"])
def test_median(entry):
    a = G.stage(entry["key"] + "/in")
    assert list(oracle.masked_median_rgb(a)) == entry["median"]
    assert [list(c) for c in oracle.edge_strip_median_colors(a)] == entry["edges"]


@pytest.mark.parametrize("name", ["squarespace", "audio_book"])
def test_bundle_median_and_fills(name):
    bg, _ = G.bundle(name)
    info = G.manifest()["bundles"][name]
    assert list(oracle.masked_median_rgb(bg)) == info["median_color"]
    assert [list(c) for c in oracle.edge_strip_median_colors(bg)] == info["edge_strip_medians"]
    for f in G.manifest()["stage"]["fill"]:
        if f.get("bundle") != name:
            continue
        size = tuple(f["size"])
        assert G.sha(oracle.fill_solid_from(bg, size)) == f["solid_sha256"]
        assert G.sha(oracle.fill_gradient_from(bg, size)) == f["gradient_sha256"]


def test_gradient_synthetic_both_axes():
    n = 0
    for f in G.manifest()["stage"]["fill"]:
        if "synthetic" not in f:
            continue
        w, h = f["size"]
        key = f"gradient/synth{f['synthetic']}/{w}x{h}"
        out = oracle.fill_gradient_from(G.stage(key + "/in"), (w, h))
        assert np.array_equal(out, G.stage(key + "/out"))
        n += 1
    assert n == 6


def test_reference_asset_weak_golden():
    # assets/draft_macro_iter_00.png: solid colour == squarespace median, text_1 pixel-exact at (27,358)
    info = G.manifest()["stage"]["asset"]
    bg, objs = G.bundle("squarespace")
    W, H = info["size"]
    canvas = oracle.fill_solid_from(bg, (W, H))
    assert list(canvas[0, 0]) == info["background_pixel"]
    obj = objs[info["object_id"]]
    x, y = info["dest"]
    out = oracle.composite(canvas, objs, [{"object_id": info["object_id"], "box": [x, y, x + obj.shape[1], y + obj.shape[0]]}])
    crop = out[y:y + obj.shape[0], x:x + obj.shape[1]]
    opaque = obj[..., 3] == 255
    assert opaque.sum() > 1000
    assert np.array_equal(crop[opaque], G.stage("asset/draft00_text1_crop")[opaque])


def test_coeffs_shape_and_normalisation():
    for in_size, out_size in [(64, 32), (33, 64), (1536, 1), (100, 99), (7, 7)]:
        k, b, ks = oracle.coeffs(in_size, out_size)
        scale = max(1.0, in_size / out_size)
        assert ks == int(np.ceil(3.0 * scale)) * 2 + 1
        assert (b[:, 0] >= 0).all() and (b[:, 0] + b[:, 1] <= in_size).all() and (b[:, 1] <= ks).all()
        assert np.abs(k.sum(axis=1) - (1 << 22)).max() <= ks  # each row sums to ~1.0 in 22-bit fixed point
