#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (it needs /root/reference and Pillow):

    python tests/golden/make_golden.py

It imports ``compositor``, ``background_resizing``, ``layout_constraints`` and the
Flex-DSL placer of ``macro_placement_test`` straight from /root/reference and records
inputs + outputs of the hot path (SURVEY.md section 8c/8d, configs C1 and C2 plus
stage-level vectors).  The fixtures travel to the GPU box; /root/reference does not.

Outputs
  bundles.npz        the two reference bundles decoded to RGBA arrays (inputs)
  stage_vectors.npz  resize / alpha_composite / premultiply / median / gradient vectors
  composites.npz     full composite() cases (C1, C2, scaled variants, semantics cases)
  manifest.json      case descriptions, placements, sha256 of every expected output,
                     Pillow / NumPy versions used
"""
from __future__ import annotations

import contextlib
import hashlib
import io
import json
import os
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)

import PIL  # noqa: E402
from PIL import Image  # noqa: E402

import background_resizing as ref_bg  # noqa: E402
import compositor as ref_comp  # noqa: E402
import layout_constraints as ref_lc  # noqa: E402

with contextlib.redirect_stdout(io.StringIO()):
    import macro_placement_test as ref_mpt  # noqa: E402


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def arr(img: Image.Image) -> np.ndarray:
    return np.array(img.convert("RGBA"), dtype=np.uint8)


def content(rng, h, w, mode):
    """Four content modes: random / binary alpha / opaque / smooth ramps with soft alpha."""
    if mode == "random":
        return rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    if mode == "binary":
        a = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        a[..., 3] = np.where(rng.random((h, w)) < 0.4, 0, 255)
        return a
    if mode == "opaque":
        a = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        a[..., 3] = 255
        return a
    yy, xx = np.mgrid[0:h, 0:w]
    a = np.zeros((h, w, 4), np.uint8)
    a[..., 0] = (xx * 255 // max(1, w - 1)).astype(np.uint8)
    a[..., 1] = (yy * 255 // max(1, h - 1)).astype(np.uint8)
    a[..., 2] = ((xx + yy) * 255 // max(1, w + h - 2)).astype(np.uint8)
    cx, cy = (w - 1) / 2.0, (h - 1) / 2.0
    r = np.sqrt(((xx - cx) / max(1.0, w / 2.0)) ** 2 + ((yy - cy) / max(1.0, h / 2.0)) ** 2)
    a[..., 3] = np.clip((1.05 - r) * 255 / 0.25, 0, 255).astype(np.uint8)
    return a


def place(tree, images, canvas_size):
    """Resolve a hand-written Flex-DSL tree with the reference placer (no VLM)."""
    placements = []
    ref_mpt._place_flex_container(tree, (0, 0), canvas_size, images, placements, "root")
    ref_mpt._clamp_boxes_to_canvas(placements, canvas_size)
    return [{"object_id": p["object_id"], "box": [int(v) for v in p["box"]]} for p in placements]


class SizeProxy:
    """The placer only reads .size (macro_placement_test.py:655,709)."""

    def __init__(self, size):
        self.size = size


def main():
    manifest = {
        "pillow": PIL.__version__,
        "numpy": np.__version__,
        "generator": "tests/golden/make_golden.py",
        "reference": "FelixMul/image_transformation (/root/reference, unmodified)",
        "cases": [],
        "stage": {},
    }
    rng = np.random.default_rng(20261018)

    # ---------------------------------------------------------------- bundles
    bundles = {}
    bundle_np = {}
    for name in ("squarespace", "audio_book"):
        base = os.path.join(REF, "output", name)
        objs = ref_comp.load_object_images(os.path.join(base, "results.json"))
        bg = ref_bg._load_background_rgba(os.path.join(base, "background.png"))
        bundles[name] = (bg, objs)
        bundle_np[f"{name}/background"] = arr(bg)
        for oid, im in objs.items():
            bundle_np[f"{name}/obj{oid}"] = arr(im)
        with open(os.path.join(base, "results.json")) as f:
            manifest.setdefault("bundles", {})[name] = {
                "results_json": json.load(f),
                "median_color": list(ref_bg._median_color_nontransparent(bg)),
                "edge_strip_medians": [list(c) for c in ref_bg._edge_strip_median_colors(bg)],
            }
    np.savez_compressed(os.path.join(HERE, "bundles.npz"), **bundle_np)

    comp_np = {}

    def add_case(name, bg_img, objs, placements, note):
        out = ref_comp.composite(bg_img, objs, placements)
        key = f"case/{name}"
        comp_np[key] = arr(out)
        return {"name": name, "note": note, "placements": placements, "sha256": sha(arr(out)),
                "canvas": list(out.size)}

    # ------------------------------------------------------- C1: squarespace 1:1
    bg, objs = bundles["squarespace"]
    with contextlib.redirect_stdout(io.StringIO()):
        size_c1 = ref_lc.compute_canvas_size((970, 250), "1:1")
    tree_c1 = {
        "direction": "column", "justify": "space-between", "align": "center", "gap_px": 12, "padding_px": 14,
        "children": [
            {"object_id": 1},
            {"object_id": 2},
            {"direction": "row", "justify": "space-between", "align": "end", "gap_px": 12, "padding_px": 0,
             "children": [{"object_id": 3}, {"object_id": 4}]},
        ],
    }
    canvas = ref_bg.fill_solid(os.path.join(REF, "output", "squarespace", "background.png"), size_c1)
    pl = place(tree_c1, objs, size_c1)
    c = add_case("c1_squarespace_1x1", canvas, objs, pl, "C1: fill_solid canvas + Flex-DSL tree, identity sizes")
    c.update(bundle="squarespace", bg="fill_solid", tree=tree_c1)
    manifest["cases"].append(c)

    # ------------------------------------------------- C2: audio_book 16:9, 9:16
    bg, objs = bundles["audio_book"]
    bgpath = os.path.join(REF, "output", "audio_book", "background.png")
    tree_169 = {
        "direction": "row", "justify": "space-around", "align": "center", "gap_px": 10, "padding_px": 8,
        "children": [
            {"object_id": 1},
            {"direction": "column", "justify": "center", "align": "start", "gap_px": 16, "padding_px": 0,
             "children": [{"object_id": 2}, {"object_id": 3}]},
        ],
    }
    tree_916 = {
        "direction": "column", "justify": "center", "align": "center", "gap_px": 20, "padding_px": 0,
        "children": [{"object_id": 1}, {"object_id": 2}, {"object_id": 3}],
    }
    for ratio, tree in (("16:9", tree_169), ("9:16", tree_916)):
        with contextlib.redirect_stdout(io.StringIO()):
            size = ref_lc.compute_canvas_size((970, 250), ratio)
        canvas = ref_bg.fill_solid(bgpath, size)
        pl = place(tree, objs, size)
        tag = ratio.replace(":", "x")
        c = add_case(f"c2_audio_book_{tag}", canvas, objs, pl, "C2: identity sizes; 9:16 has an overhanging box")
        c.update(bundle="audio_book", bg="fill_solid", tree=tree)
        manifest["cases"].append(c)
        # scaled variants: same tree resolved with size-proxy objects
        for s in (0.5, 0.75, 1.5):
            proxies = {oid: SizeProxy((max(1, round(im.size[0] * s)), max(1, round(im.size[1] * s))))
                       for oid, im in objs.items()}
            pl = place(tree, proxies, size)
            c = add_case(f"c2_audio_book_{tag}_s{s}", canvas, objs, pl, f"C2 scaled variant x{s}: LANCZOS resample path")
            c.update(bundle="audio_book", bg="fill_solid", tree=tree, scale=s)
            manifest["cases"].append(c)
        # gradient canvas + over (background_resizing.fill_gradient has no caller; keep parity)
    # C1 scaled too (down and anisotropic boxes straight into composite())
    bg, objs = bundles["squarespace"]
    canvas = ref_bg.fill_solid(os.path.join(REF, "output", "squarespace", "background.png"), size_c1)
    pl = [
        {"object_id": 2, "box": [-40, -25, 300, 170]},
        {"object_id": 1, "box": [100.9, 60.2, 420.5, 131.7]},
        {"object_id": "3", "box": [30, 200, 480, 480]},
        {"object_id": 4, "box": [350, 420, 520, 500]},
        {"object_id": 99, "box": [0, 0, 10, 10]},
        {"object_id": 4, "box": [200, 300, 200, 290]},
    ]
    c = add_case("c1_squarespace_scaled_mixed", canvas, objs, pl,
                 "float boxes, str id, unknown id, degenerate box, negative dest, overhang, up+down scale")
    c.update(bundle="squarespace", bg="fill_solid")
    manifest["cases"].append(c)

    # bundle background (partially transparent canvas, da in {0,255}) as the canvas itself
    bgs, objs = bundles["squarespace"]
    pl = [{"object_id": 2, "box": [10, 10, 10 + 357, 10 + 207]}, {"object_id": 3, "box": [500, 40, 900, 253]}]
    c = add_case("squarespace_on_transparent_background", bgs, objs, pl, "dst alpha 0/255 canvas, clipping at bottom")
    c.update(bundle="squarespace", bg="bundle_background")
    manifest["cases"].append(c)

    # ------------------------------------------ reference's own known-answer test
    bgk = Image.new("RGBA", (10, 10), (255, 0, 0, 255))
    objk = Image.new("RGBA", (2, 2), (0, 255, 0, 255))
    outk = ref_comp.composite(bgk, {1: objk}, [{"object_id": 1, "box": [4, 4, 6, 6]}])
    assert outk.getpixel((4, 4))[:3] == (0, 255, 0)
    comp_np["case/reference_known_answer"] = arr(outk)
    manifest["cases"].append({"name": "reference_known_answer", "note": "tests/test_compositor.py:5-11",
                              "sha256": sha(arr(outk)), "canvas": [10, 10], "synthetic": "known_answer",
                              "placements": [{"object_id": 1, "box": [4, 4, 6, 6]}]})

    # ------------------------------------------------ synthetic semantics cases
    def synth_case(name, W, H, bg_mode, obj_shapes, placements, note, seed):
        r = np.random.default_rng(seed)
        bga = content(r, H, W, bg_mode)
        objs_a = {i + 1: content(r, sh, sw, m) for i, (sw, sh, m) in enumerate(obj_shapes)}
        out = ref_comp.composite(Image.fromarray(bga, "RGBA"),
                                 {k: Image.fromarray(v, "RGBA") for k, v in objs_a.items()}, placements)
        comp_np[f"case/{name}"] = arr(out)
        comp_np[f"in/{name}/bg"] = bga
        for k, v in objs_a.items():
            comp_np[f"in/{name}/obj{k}"] = v
        manifest["cases"].append({"name": name, "note": note, "placements": placements, "sha256": sha(arr(out)),
                                  "canvas": [W, H], "synthetic": "stored_inputs"})

    synth_case("synth_overlap_zorder", 200, 150, "opaque",
               [(90, 70, "smooth"), (64, 64, "random"), (120, 40, "binary")],
               [{"object_id": 1, "box": [10, 10, 130, 100]}, {"object_id": 2, "box": [60, 40, 110, 120]},
                {"object_id": 3, "box": [30, 60, 190, 110]}, {"object_id": 1, "box": [100, 5, 160, 60]}],
               "3-deep overlap, later on top, same cutout placed twice", 11)
    synth_case("synth_general_dst_alpha", 160, 120, "random",
               [(50, 40, "random"), (33, 77, "smooth")],
               [{"object_id": 1, "box": [5, 5, 100, 90]}, {"object_id": 2, "box": [70, 20, 150, 118]},
                {"object_id": 1, "box": [20, 60, 70, 100]}],
               "random dst alpha exercises the general over (division) path", 12)
    synth_case("synth_clipping", 96, 64, "opaque",
               [(40, 30, "smooth"), (20, 20, "random")],
               [{"object_id": 1, "box": [-30, -20, 50, 40]}, {"object_id": 2, "box": [80, 50, 130, 100]},
                {"object_id": 1, "box": [-10, -10, 110, 80]}, {"object_id": 2, "box": [200, 200, 230, 230]},
                {"object_id": 2, "box": [-50, 10, -20, 40]}],
               "negative dest, overhang, larger than canvas, fully outside", 13)
    synth_case("synth_extreme_scales", 128, 128, "opaque",
               [(300, 5, "random"), (3, 400, "random"), (1, 1, "random"), (257, 129, "smooth")],
               [{"object_id": 1, "box": [4, 4, 24, 9]}, {"object_id": 2, "box": [40, 2, 49, 4]},
                {"object_id": 2, "box": [60, 2, 62, 120]}, {"object_id": 3, "box": [70, 70, 100, 90]},
                {"object_id": 4, "box": [5, 30, 18, 126]}, {"object_id": 4, "box": [30, 100, 120, 104]}],
               "15x downscale, Pillow-12 tall-image vertical-first (3x400 -> 9x2), 1x1 upscale, single-axis", 14)
    synth_case("synth_single_axis", 100, 100, "opaque",
               [(40, 30, "random")],
               [{"object_id": 1, "box": [5, 5, 45, 80]}, {"object_id": 1, "box": [50, 10, 95, 40]},
                {"object_id": 1, "box": [55, 60, 95, 90]}],
               "height-only, width-only and identity placements of the same cutout", 15)
    np.savez_compressed(os.path.join(HERE, "composites.npz"), **comp_np)

    # ---------------------------------------------------------- stage vectors
    st = {}
    shapes = [  # (sw, sh, w, h)
        (64, 48, 32, 24), (64, 48, 47, 31), (33, 21, 64, 48), (1, 1, 4, 4), (64, 48, 7, 5), (5, 7, 64, 48),
        (100, 3, 9, 3), (120, 90, 13, 200), (50, 50, 50, 20), (50, 50, 20, 50), (3, 400, 9, 2), (3, 400, 2, 399),
        (97, 61, 96, 60), (16, 16, 1, 1), (200, 10, 11, 10),
    ]
    res_list = []
    for i, (sw, sh, w, h) in enumerate(shapes):
        for mode in ("random", "binary", "opaque", "smooth"):
            src = content(rng, sh, sw, mode)
            out = arr(Image.fromarray(src, "RGBA").resize((w, h), Image.LANCZOS))
            key = f"resize/{i}_{mode}"
            st[key + "/src"] = src
            st[key + "/out"] = out
            res_list.append({"key": key, "src": [sw, sh], "dst": [w, h], "mode": mode, "sha256": sha(out)})
    manifest["stage"]["resize"] = res_list

    # premultiply / unpremultiply: exhaustive (c, a) grid
    cc, aa = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8))
    grid = np.stack([cc, cc[::-1], (cc.astype(int) * 7 % 256).astype(np.uint8), aa], axis=-1)
    pm = np.array(Image.fromarray(grid, "RGBA").convert("RGBa"))
    # PIL arrays of mode RGBa export the raw premultiplied bytes
    st["premul/in"] = grid
    st["premul/out"] = pm.astype(np.uint8)
    valid = np.minimum(grid[..., :3], grid[..., 3:4])  # premultiplied input must satisfy c <= a
    pin = np.concatenate([valid, grid[..., 3:4]], axis=-1).astype(np.uint8)
    im = Image.frombuffer("RGBa", (256, 256), pin.tobytes(), "raw", "RGBa", 0, 1)
    st["unpremul/in"] = pin
    st["unpremul/out"] = np.array(im.convert("RGBA"), dtype=np.uint8)
    manifest["stage"]["premul"] = {"sha256": sha(st["premul/out"])}
    manifest["stage"]["unpremul"] = {"sha256": sha(st["unpremul/out"])}

    # alpha_composite: exhaustive (sa, da) grids for a few colour pairs + random pairs
    over_list = []
    sa, da = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8))
    for j, (sc, dc) in enumerate([((255, 0, 128), (0, 255, 64)), ((13, 200, 77), (250, 3, 190)),
                                  ((0, 0, 0), (255, 255, 255)), ((255, 255, 255), (0, 0, 0))]):
        s = np.zeros((256, 256, 4), np.uint8)
        d = np.zeros((256, 256, 4), np.uint8)
        s[..., :3] = sc
        d[..., :3] = dc
        s[..., 3] = sa
        d[..., 3] = da
        out = arr(Image.alpha_composite(Image.fromarray(d, "RGBA"), Image.fromarray(s, "RGBA")))
        st[f"over/grid{j}/out"] = out
        over_list.append({"key": f"over/grid{j}", "src_rgb": sc, "dst_rgb": dc, "sha256": sha(out)})
    for j in range(4):
        s = content(rng, 61, 83, "random")
        d = content(rng, 61, 83, "random" if j % 2 == 0 else "opaque")
        out = arr(Image.alpha_composite(Image.fromarray(d, "RGBA"), Image.fromarray(s, "RGBA")))
        st[f"over/rand{j}/src"] = s
        st[f"over/rand{j}/dst"] = d
        st[f"over/rand{j}/out"] = out
        over_list.append({"key": f"over/rand{j}", "sha256": sha(out)})
    manifest["stage"]["over"] = over_list

    # medians / fills (background_resizing.py)
    med_list = []
    for j, (h, w, mode) in enumerate([(1, 1, "random"), (1, 2, "random"), (7, 1, "binary"), (40, 25, "binary"),
                                      (77, 13, "random"), (31, 33, "opaque"), (64, 64, "smooth")]):
        a = content(rng, h, w, mode)
        if j == 4:
            a[..., 3] = 0  # fully transparent -> fallback branch (background_resizing.py:15-19)
        imj = Image.fromarray(a, "RGBA")
        st[f"median/{j}/in"] = a
        med_list.append({"key": f"median/{j}", "median": list(ref_bg._median_color_nontransparent(imj)),
                         "edges": [list(c) for c in ref_bg._edge_strip_median_colors(imj)]})
    manifest["stage"]["median"] = med_list

    fill_list = []
    for name in ("squarespace", "audio_book"):
        path = os.path.join(REF, "output", name, "background.png")
        for size in ((492, 492), (657, 369), (369, 657), (1, 1), (2, 5)):
            solid = arr(ref_bg.fill_solid(path, size))
            grad = arr(ref_bg.fill_gradient(path, size))
            fill_list.append({"bundle": name, "size": list(size), "solid_sha256": sha(solid),
                              "solid_px": [int(v) for v in solid[0, 0]], "gradient_sha256": sha(grad),
                              "gradient_first_row": grad[0].tolist() if size[0] <= 8 else None,
                              "gradient_row0_sha256": sha(grad[0]), "gradient_col0_sha256": sha(grad[:, 0])})
            if size == (657, 369):
                st[f"gradient/{name}/657x369/row0"] = grad[0]
                st[f"gradient/{name}/657x369/col0"] = grad[:, 0]
    # a synthetic background whose edges force each gradient direction
    for j, horizontal in enumerate((True, False)):
        a = np.zeros((40, 60, 4), np.uint8)
        a[..., 3] = 255
        if horizontal:
            a[..., :3] = 120
            a[:, :30, 0] = 10
            a[:, 30:, 0] = 240
            a[:20, :, 1] = 119
        else:
            a[..., :3] = 120
            a[:20, :, 2] = 5
            a[20:, :, 2] = 250
        a[5:9, 7:11, 3] = 0
        tmp = os.path.join("/tmp", f"golden_bg_{j}.png")
        Image.fromarray(a, "RGBA").save(tmp)
        for size in ((97, 41), (1, 9), (300, 2)):
            grad = arr(ref_bg.fill_gradient(tmp, size))
            st[f"gradient/synth{j}/{size[0]}x{size[1]}/in"] = a
            st[f"gradient/synth{j}/{size[0]}x{size[1]}/out"] = grad
            fill_list.append({"synthetic": j, "size": list(size), "gradient_sha256": sha(grad)})
    manifest["stage"]["fill"] = fill_list

    # weak golden shipped by the reference: assets/draft_macro_iter_00.png shows text_1.png
    # pixel-exact at (27,358) on the fill_solid colour
    asset = arr(Image.open(os.path.join(REF, "assets", "draft_macro_iter_00.png")))
    tw, th = bundles["squarespace"][1][3].size
    crop = asset[358:358 + th, 27:27 + tw]
    st["asset/draft00_text1_crop"] = crop
    manifest["stage"]["asset"] = {"file": "assets/draft_macro_iter_00.png", "size": [asset.shape[1], asset.shape[0]],
                                  "object_id": 3, "dest": [27, 358],
                                  "background_pixel": [int(v) for v in asset[0, 0]]}

    np.savez_compressed(os.path.join(HERE, "stage_vectors.npz"), **st)
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, default=lambda o: list(o))
    for fn in ("bundles.npz", "stage_vectors.npz", "composites.npz", "manifest.json"):
        print(fn, os.path.getsize(os.path.join(HERE, fn)) // 1024, "KiB")


if __name__ == "__main__":
    main()
