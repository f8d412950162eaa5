#!/usr/bin/env python3
"""Golden fixtures for the SURVEY 8(f) rows (contact sheet, candidates grid, API downscale), generated from the
UNMODIFIED reference in the build container:

    python tests/golden/make_golden_sheets.py      ->  tests/golden/sheets.npz, sheets_manifest.json

  contact/<bundle>      macro_placement_test._build_labeled_contact_sheet (:162-242) on the bundle as shipped
  grid                  macro_placement_test._compose_candidates_grid (:1332-1345) on four drafts of different sizes
  api/<name>            the RGB downscale of api_client._prepare_image_b64_for_api (:97-108), before its JPEG encode
  flow/<case>           layout_constraints.pack_flow (:273-327) boxes of scaled objects -> fill_solid -> composite
"""
from __future__ import annotations

import contextlib
import hashlib
import io
import json
import os
import sys
import tempfile
from pathlib import Path

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)

import PIL  # noqa: E402
from PIL import Image  # noqa: E402

with contextlib.redirect_stdout(io.StringIO()):
    import macro_placement_test as ref_mpt  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    out, man = {}, {"pillow": PIL.__version__, "contact": {}, "grid": {}, "api": {}}
    for bundle in ("squarespace", "audio_book"):
        rj = f"{REF}/output/{bundle}/results.json"
        sheet = ref_mpt._build_labeled_contact_sheet(f"{REF}/output/{bundle}/objects", rj)
        a = np.array(sheet.convert("RGBA"), dtype=np.uint8)
        out[f"contact/{bundle}"] = a
        items = sorted(json.load(open(rj)), key=lambda it: int(it["object_id"]))
        man["contact"][bundle] = {"sha256": sha(a), "size": list(sheet.size),
                                  "items": [{"object_id": int(it["object_id"]), "label": it.get("label"),
                                             "filename": it["filename"]} for it in items]}
    # candidates grid: four drafts of different sizes (the first sets the cell size)
    rng = np.random.default_rng(7)
    d0 = Image.open(f"{REF}/assets/draft_macro_iter_00.png").convert("RGBA")
    d1 = Image.open(f"{REF}/assets/draft_macro_iter_01.png").convert("RGBA")
    d2 = Image.open(f"{REF}/output/squarespace/background.png").convert("RGBA")       # 970x250, binary alpha holes
    soft = rng.integers(0, 256, (300, 200, 4), dtype=np.uint8)                          # random soft alpha, upscaled
    d3 = Image.fromarray(soft, "RGBA")
    drafts = [d0, d1, d2, d3]
    with tempfile.TemporaryDirectory() as td:
        paths = []
        for i, d in enumerate(drafts):
            p = Path(td) / f"draft_{i}.png"
            d.save(p)
            paths.append(p)
        outp = Path(td) / "grid.png"
        ref_mpt._compose_candidates_grid(paths, outp)
        grid = np.array(Image.open(outp).convert("RGBA"), dtype=np.uint8)
    for i, d in enumerate(drafts):
        out[f"grid/in{i}"] = np.array(d, dtype=np.uint8)
    out["grid/out"] = grid
    man["grid"] = {"sha256": sha(grid), "n": len(drafts)}
    # API downscale (api_client.py:102-108)
    for name in ("squarespace", "audio_book"):
        im = Image.open(f"{REF}/input/{name}.jpg").convert("RGB")
        for max_side in (512, 200):
            w, h = im.size
            scale = max_side / float(max(w, h))
            small = im.resize((max(1, int(w * scale)), max(1, int(h * scale))), Image.LANCZOS)
            out[f"api/{name}/{max_side}"] = np.array(small, dtype=np.uint8)
            man["api"][f"{name}/{max_side}"] = {"sha256": sha(np.array(small)), "size": list(small.size)}
        out[f"api/{name}/in"] = np.array(im, dtype=np.uint8)
    # pack_flow (layout_constraints.py:273-327): the in-tree producer of boxes whose size differs from the cutout's
    import compositor as ref_comp
    import background_resizing as ref_bg
    import layout_constraints as ref_lc

    man["flow"] = {}
    for bundle, ratio, scales, params in (
            ("squarespace", "9:16", {1: 1.3, 2: 0.8, 3: 0.66, 4: 2.0}, {"align": "center", "orientation": "auto"}),
            ("squarespace", "16:9", {1: 0.5, 2: 0.45, 3: 0.7, 4: 1.0}, {"align": "left", "orientation": "horizontal",
                                                                       "global_margin_px": 7, "global_spacing_px": 11}),
            ("audio_book", "1:1", {1: 0.9, 2: 0.75, 3: 1.5}, {"align": "center", "orientation": "vertical",
                                                             "global_spacing_px": 3})):
        rj = f"{REF}/output/{bundle}/results.json"
        objs = ref_comp.load_object_images(rj)
        items = json.load(open(rj))
        meta = {int(it["object_id"]): ref_lc.ObjectMeta(int(it["object_id"]), it.get("label", ""), it["filename"],
                                                        objs[int(it["object_id"])].width, objs[int(it["object_id"])].height)
                for it in items}
        scaled = [ref_lc.ObjectMeta(m.object_id, m.label, m.file, max(1, int(m.width * scales[m.object_id])),
                                    max(1, int(m.height * scales[m.object_id]))) for m in meta.values()]
        canvas_size = ref_lc.compute_canvas_size((970, 250), ratio)
        pls, _ = ref_lc.pack_flow(scaled, canvas_size, params, meta)
        layout = ref_lc.layout_final_json(pls, canvas_size, 0.0, params["align"])
        bg = ref_bg.fill_solid(f"{REF}/output/{bundle}/background.png", canvas_size)
        res = ref_comp.composite(bg, objs, layout["placements"])
        key = f"{bundle}_{ratio.replace(':', 'x')}"
        out[f"flow/{key}"] = np.array(res, dtype=np.uint8)
        man["flow"][key] = {"bundle": bundle, "canvas": list(canvas_size), "placements": layout["placements"],
                            "sha256": sha(np.array(res))}
    np.savez_compressed(os.path.join(HERE, "sheets.npz"), **out)
    json.dump(man, open(os.path.join(HERE, "sheets_manifest.json"), "w"), indent=1)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
