"""contact sheet, candidates grid and the API downscale (image_transformation_b200/sheets.py)
against golden outputs of the unmodified reference functions (tests/golden/sheets.npz, made by
make_golden_sheets.py).  CPU tests pin the oracle and the host size rules on those goldens; the GPU tests run the
CUDA path through the mirrored functions, bit-exact."""
import json
import os

import numpy as np
import pytest
from PIL import Image

import golden_io as G
import oracle
from image_transformation_b200 import sheets

Z = np.load(os.path.join(G.GOLDEN, "sheets.npz"))
MAN = json.load(open(os.path.join(G.GOLDEN, "sheets_manifest.json")))
THUMB, LABEL_H, COLS = (256, 256), 72, 4


def _thumb_rows(a: np.ndarray) -> np.ndarray:
    """The sheet without its label strips (text rendering is host PIL/FreeType in the reference and here)."""
    cell_h = THUMB[1] + LABEL_H
    rows = [a[r * cell_h: r * cell_h + THUMB[1]] for r in range(a.shape[0] // cell_h)]
    return np.concatenate(rows, axis=0)


def _sheet_inputs(bundle):
    _, objs = G.bundle(bundle)
    items = MAN["contact"][bundle]["items"]
    images = [objs[it["object_id"]] for it in items]
    placements = []
    for idx, im in enumerate(images):
        h, w = im.shape[:2]
        tw, th = sheets.thumbnail_size((w, h), THUMB) or (w, h)
        x = (idx % COLS) * THUMB[0] + (THUMB[0] - tw) // 2
        y = (idx // COLS) * (THUMB[1] + LABEL_H) + (THUMB[1] - th) // 2
        placements.append({"object_id": idx, "box": [x, y, x + tw, y + th]})
    return images, placements, items


# ------------------------------------------------------------------------------ CPU: oracle + host rules
def test_thumbnail_size_matches_pillow():
    rng = np.random.default_rng(5)
    for _ in range(1500):
        w, h = (int(v) for v in rng.integers(1, 700, 2))
        b = tuple(int(v) for v in rng.integers(1, 400, 2))
        im = Image.new("L", (w, h))
        im.thumbnail(b)
        assert (sheets.thumbnail_size((w, h), b) or (w, h)) == im.size, ((w, h), b)


def test_api_downscale_size_rule():
    assert sheets.api_downscale_size((970, 250), 512) == (512, 131)
    assert sheets.api_downscale_size((970, 250), 200) == (200, 51)
    assert sheets.api_downscale_size((300, 512), 512) is None
    assert sheets.api_downscale_size((3, 2000), 512) == (1, 512)


@pytest.mark.parametrize("bundle", ["squarespace", "audio_book"])
def test_oracle_contact_sheet_matches_reference(bundle):
    images, placements, _ = _sheet_inputs(bundle)
    exp = Z[f"contact/{bundle}"]
    assert G.sha(exp) == MAN["contact"][bundle]["sha256"]
    bg = np.full(exp.shape, 255, np.uint8)
    got = oracle.composite(bg, dict(enumerate(images)), placements)
    assert np.array_equal(_thumb_rows(got), _thumb_rows(exp))


def test_oracle_candidates_grid_matches_reference():
    ins = [Z[f"grid/in{i}"] for i in range(MAN["grid"]["n"])]
    exp = Z["grid/out"]
    assert G.sha(exp) == MAN["grid"]["sha256"]
    h, w = ins[0].shape[:2]
    bg = np.full((2 * h, 2 * w, 4), 255, np.uint8)
    pl = [{"object_id": i, "box": [x, y, x + w, y + h]} for i, (x, y) in enumerate([(0, 0), (w, 0), (0, h), (w, h)])]
    assert np.array_equal(oracle.composite(bg, dict(enumerate(ins)), pl), exp)


@pytest.mark.parametrize("key", sorted(MAN["api"]))
def test_oracle_rgb_downscale_matches_reference(key):
    name, _ = key.split("/")
    src = Z[f"api/{name}/in"]
    exp = Z[f"api/{key}"]
    rgba = np.dstack([src, np.full(src.shape[:2], 255, np.uint8)])
    got = oracle.resize_rgba_lanczos(rgba, tuple(MAN["api"][key]["size"]))
    assert np.array_equal(got[..., :3], exp) and (got[..., 3] == 255).all()


@pytest.mark.parametrize("key", sorted(MAN["flow"]))
def test_oracle_pack_flow_composite_matches_reference(key):
    c = MAN["flow"][key]
    _, objs = G.bundle(c["bundle"])
    W, H = c["canvas"]
    bg = np.zeros((H, W, 4), np.uint8)
    bg[...] = tuple(G.manifest()["bundles"][c["bundle"]]["median_color"]) + (255,)
    exp = Z[f"flow/{key}"]
    assert G.sha(exp) == c["sha256"]
    assert np.array_equal(oracle.composite(bg, objs, c["placements"]), exp)


# ------------------------------------------------------------------------------ GPU: the mirrored functions
@pytest.mark.gpu
@pytest.mark.parametrize("key", sorted(MAN["flow"]))
def test_pack_flow_layouts_through_dropin(key):
    """Boxes from the reference's pack_flow (scaled objects: the in-tree producer of non-identity boxes)."""
    from image_transformation_b200 import compositor

    c = MAN["flow"][key]
    _, objs = G.bundle(c["bundle"])
    W, H = c["canvas"]
    bg = Image.new("RGBA", (W, H), tuple(G.manifest()["bundles"][c["bundle"]]["median_color"]) + (255,))
    out = compositor.composite(bg, {k: Image.fromarray(v, "RGBA") for k, v in objs.items()}, c["placements"])
    assert np.array_equal(np.asarray(out), Z[f"flow/{key}"])


@pytest.mark.gpu
def test_agentic_native_size_loop():
    _, objs = G.bundle("audio_book")
    imgs = {k: Image.fromarray(v, "RGBA") for k, v in objs.items()}
    bg = Image.new("RGBA", (657, 369), (38, 73, 115, 255))
    pl = [{"object_id": k, "box": [20 + 90 * i, 10 + 30 * i, 20 + 90 * i + im.width, 10 + 30 * i + im.height]}
          for i, (k, im) in enumerate(sorted(imgs.items()))]
    exp = bg.copy()
    for p in pl:
        exp.alpha_composite(imgs[p["object_id"]], dest=(p["box"][0], p["box"][1]))
    assert np.array_equal(np.asarray(sheets.composite_native_size(bg, imgs, pl)), np.asarray(exp))
    pl[1]["box"][2] += 1
    with pytest.raises(ValueError, match="Placement size mismatch"):
        sheets.composite_native_size(bg, imgs, pl)

@pytest.mark.gpu
@pytest.mark.parametrize("bundle", ["squarespace", "audio_book"])
def test_build_labeled_contact_sheet(bundle, tmp_path):
    images, _, items = _sheet_inputs(bundle)
    for it, a in zip(items, images):
        p = tmp_path / it["filename"]
        p.parent.mkdir(parents=True, exist_ok=True)
        Image.fromarray(a, "RGBA").save(p)
    rj = tmp_path / "results.json"
    rj.write_text(json.dumps(items))
    sheet = sheets.build_labeled_contact_sheet(str(tmp_path / "objects"), str(rj))
    exp = Z[f"contact/{bundle}"]
    assert sheet.mode == "RGBA" and sheet.size == tuple(MAN["contact"][bundle]["size"])
    got = np.asarray(sheet)
    assert np.array_equal(_thumb_rows(got), _thumb_rows(exp))
    # labels are host-side PIL text in both: identical whenever the same font is installed
    diff = (got != exp).any(axis=-1)
    cell_h = THUMB[1] + LABEL_H
    assert not diff.reshape(-1, cell_h, diff.shape[1])[:, :THUMB[1]].any()


@pytest.mark.gpu
def test_compose_candidates_grid(tmp_path):
    paths = []
    for i in range(MAN["grid"]["n"]):
        p = tmp_path / f"draft_{i}.png"
        Image.fromarray(Z[f"grid/in{i}"], "RGBA").save(p)
        paths.append(p)
    paths.append(tmp_path / "missing.png")  # skipped, as in the reference
    out = tmp_path / "grid.png"
    sheets.compose_candidates_grid(paths, out)
    assert np.array_equal(np.asarray(Image.open(out).convert("RGBA")), Z["grid/out"])


@pytest.mark.gpu
@pytest.mark.parametrize("key", sorted(MAN["api"]))
def test_prepare_image_for_api(key, tmp_path):
    name, max_side = key.split("/")
    p = tmp_path / f"{name}.png"  # lossless container for the decoded reference input
    Image.fromarray(Z[f"api/{name}/in"], "RGB").save(p)
    got = sheets.prepare_image_for_api(p, int(max_side))
    assert got.mode == "RGB" and np.array_equal(np.asarray(got), Z[f"api/{key}"])
    assert isinstance(sheets.prepare_image_b64_for_api(p, int(max_side)), str)


@pytest.mark.gpu
def test_thumbnail_and_resize_helpers_vs_pillow():
    rng = np.random.default_rng(9)
    for (w, h) in [(447, 116), (131, 32), (1536, 900), (90, 700), (3, 400)]:
        a = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        im = Image.fromarray(a, "RGBA")
        ref = im.copy()
        ref.thumbnail((256, 256), Image.LANCZOS)
        assert np.array_equal(np.asarray(sheets.thumbnail_rgba(im, (256, 256))), np.asarray(ref)), (w, h)
        rgb = im.convert("RGB")
        assert np.array_equal(np.asarray(sheets.resize_rgb_lanczos(rgb, (w // 2 + 1, h // 3 + 1))),
                              np.asarray(rgb.resize((w // 2 + 1, h // 3 + 1), Image.LANCZOS)))
    with pytest.raises(ValueError):
        sheets.resize_rgb_lanczos(Image.new("RGBA", (4, 4)), (2, 2))
