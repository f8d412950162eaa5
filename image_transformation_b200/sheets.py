"""Raster helpers either side of the compositor (the "next" rows): the same
two operations -- LANCZOS resize and integer alpha-over -- as the reference uses them outside
``composite()``.  Everything here is host glue over the CUDA path (``compositor.composite`` and
the stand-alone resampler); text labels stay host-side PIL (ImageDraw/FreeType), exactly as in the
reference.

Reference interfaces mirrored (file:line in /root/reference):
  * ``_build_labeled_contact_sheet``  macro_placement_test.py:162-242
        thumbnail((256, 256), LANCZOS) of every cutout + alpha_composite onto a white sheet + labels
  * ``_compose_candidates_grid``      macro_placement_test.py:1332-1345
        up to four drafts resized to the first one's size, 2 x 2 grid on white
  * ``_prepare_image_b64_for_api``    api_client.py:97-112   (the RGB LANCZOS downscale; JPEG/base64 stay host)
  * the agentic compositor node's raster loop  agentic/nodes/compositor.py:36-43  (native-size overlays only)

``Image.thumbnail`` on a plain RGBA image is a LANCZOS ``resize`` to the aspect-preserving size
(PIL Image.py: ``preserve_aspect_ratio``; ``draft`` only applies to JPEG files and the RGBA branch of
``resize`` ignores ``reducing_gap``), so thumbnail + alpha_composite at an offset is one placement of
``composite()``: the contact sheet and the candidates grid are single fused launches.
An RGB resize is the RGBA resize with alpha 255 (premultiply and un-premultiply are the identity
for alpha 255 and Pillow resamples every band with the same 8-bit fixed-point code).
"""
from __future__ import annotations

import base64
import io
import json
import math
from pathlib import Path
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
from PIL import Image, ImageDraw, ImageFont

from . import compositor as _compositor


# ---------------------------------------------------------------------------- size rules
def thumbnail_size(size: Tuple[int, int], bound: Tuple[float, float]) -> Optional[Tuple[int, int]]:
    """Size ``Image.thumbnail(bound)`` gives an image of ``size``; None when it is left alone
    (thumbnail never enlarges).  The side the bound does not limit is ``other * aspect`` rounded
    to whichever neighbouring integer reproduces the aspect ratio better (down on a tie), at least 1."""
    width, height = size
    bx, by = math.floor(bound[0]), math.floor(bound[1])
    if bx >= width and by >= height:
        return None
    aspect = width / height

    def nearest(number: float, err) -> int:
        lo, hi = math.floor(number), math.ceil(number)
        return max(lo if err(lo) <= err(hi) else hi, 1)

    if bx / by >= aspect:
        return nearest(by * aspect, lambda n: abs(aspect - n / by)), by
    return bx, nearest(bx / aspect, lambda n: 0 if n == 0 else abs(aspect - bx / n))


def api_downscale_size(size: Tuple[int, int], max_side: int = 512) -> Optional[Tuple[int, int]]:
    """api_client.py:104-108: longest side to <= max_side, truncating; None if already small enough."""
    w, h = size
    if max(w, h) <= max_side:
        return None
    scale = max_side / float(max(w, h))
    return max(1, int(w * scale)), max(1, int(h * scale))


# ---------------------------------------------------------------------------- resizes
def resize_rgba_lanczos(img: Image.Image, size: Tuple[int, int]) -> Image.Image:
    """``img.resize(size, Image.LANCZOS)`` for an RGBA image, on the GPU (stand-alone resampler: unlike an over
    onto a transparent canvas it keeps the colours of pixels whose alpha resampled to 0)."""
    if img.mode != "RGBA":
        raise ValueError("image has wrong mode")
    w, h = int(size[0]), int(size[1])
    if (w, h) == img.size:
        return img.copy()
    import torch

    from . import _native, batch as _batch

    _native.require_gpu()
    src = torch.from_numpy(np.array(img, dtype=np.uint8)).cuda()  # np.array: a writable copy
    # Pillow >= 12 runs the vertical pass first on very tall images (PIL Image.py:2431-2435)
    tall = _compositor.TALL_IMAGE_VERTICAL_FIRST and img.height > 100 * img.width and h < img.height
    out = _batch.resize_rgba_lanczos(src, (w, h), vertical_first=tall)
    return Image.fromarray(out.cpu().numpy(), "RGBA")


def resize_rgb_lanczos(img: Image.Image, size: Tuple[int, int]) -> Image.Image:
    """``img.resize(size, Image.LANCZOS)`` for an RGB image, on the GPU (alpha 255 rides along)."""
    if img.mode != "RGB":
        raise ValueError("image has wrong mode")
    w, h = int(size[0]), int(size[1])
    if (w, h) == img.size:
        return img.copy()
    rgba = np.empty((img.height, img.width, 4), np.uint8)
    rgba[..., :3] = np.asarray(img, dtype=np.uint8)
    rgba[..., 3] = 255
    out = np.asarray(resize_rgba_lanczos(Image.fromarray(rgba, "RGBA"), (w, h)))
    return Image.fromarray(np.ascontiguousarray(out[..., :3]), "RGB")


def thumbnail_rgba(img: Image.Image, bound: Tuple[int, int] = (256, 256)) -> Image.Image:
    """Copy of ``img`` after ``thumbnail(bound, Image.LANCZOS)``."""
    final = thumbnail_size(img.size, bound)
    if final is None or final == img.size:
        return img.copy()
    return resize_rgba_lanczos(img, final)


def prepare_image_for_api(image_path: Union[str, Path], max_side: int = 512) -> Image.Image:
    """The raster part of api_client.py:97-108: open, convert to RGB, LANCZOS-downscale the longest side."""
    im = Image.open(Path(image_path)).convert("RGB")
    target = api_downscale_size(im.size, max_side)
    return im if target is None else resize_rgb_lanczos(im, target)


def prepare_image_b64_for_api(image_path: Union[str, Path], max_side: int = 512) -> str:
    """api_client.py:97-112 with the downscale on the GPU (JPEG encode and base64 stay on the host)."""
    buf = io.BytesIO()
    prepare_image_for_api(image_path, max_side).save(buf, format="JPEG", quality=85)
    buf.seek(0)
    return base64.b64encode(buf.read()).decode("utf-8")


# ---------------------------------------------------------------------------- contact sheet
def _load_font(font_size: int):
    # same fallback chain as macro_placement_test.py:175-187
    for name in ("DejaVuSans.ttf", "/usr/share/fonts/truetype/dejavu/DejaVuSans.ttf"):
        try:
            return ImageFont.truetype(name, size=font_size)
        except Exception:
            pass
    try:
        return ImageFont.load_default()
    except Exception:
        return None


def contact_sheet(images: Sequence[Image.Image], labels: Optional[Sequence[str]] = None,
                  thumb_size: Tuple[int, int] = (256, 256), cols: int = 4, label_height: int = 72,
                  font_size: int = 24) -> Image.Image:
    """White sheet of centred LANCZOS thumbnails, ``cols`` per row, each over a ``label_height`` strip
    (labels drawn on the host when given).  The raster is ONE ``composite()`` launch."""
    if not images:
        return Image.new("RGBA", (thumb_size[0], thumb_size[1] + label_height), (255, 255, 255, 255))
    rows = (len(images) + cols - 1) // cols
    cell_w, cell_h = thumb_size[0], thumb_size[1] + label_height
    sheet = Image.new("RGBA", (cols * cell_w, rows * cell_h), (255, 255, 255, 255))
    placements, objects = [], {}
    for idx, im in enumerate(images):
        tw, th = thumbnail_size(im.size, thumb_size) or im.size
        x = (idx % cols) * cell_w + (cell_w - tw) // 2
        y = (idx // cols) * cell_h + (thumb_size[1] - th) // 2
        objects[idx] = im
        placements.append({"object_id": idx, "box": [x, y, x + tw, y + th]})
    sheet = _compositor.composite(sheet, objects, placements)
    if labels is not None:
        draw = ImageDraw.Draw(sheet)
        font = _load_font(font_size)
        for idx, label in enumerate(labels):
            x_cell, y_cell = (idx % cols) * cell_w, (idx // cols) * cell_h
            try:
                bbox = draw.textbbox((0, 0), label, font=font)
                tw, th_text = bbox[2] - bbox[0], bbox[3] - bbox[1]
            except Exception:
                tw, th_text = int(len(label) * 7), 12
            tx = x_cell + (cell_w - tw) // 2
            ty = y_cell + thumb_size[1] + max(0, (label_height - th_text) // 2)
            draw.text((tx, ty), label, fill=(0, 0, 0, 255), font=font)
    return sheet


def build_labeled_contact_sheet(objects_dir: str, results_json_path: str, thumb_size: Tuple[int, int] = (256, 256),
                                cols: int = 4, label_height: int = 72, font_size: int = 24) -> Image.Image:
    """Drop-in for ``_build_labeled_contact_sheet`` (macro_placement_test.py:162-242): same arguments
    (``objects_dir`` is unused there as well), items sorted by object id, label falls back to ``id_<n>``."""
    with open(results_json_path, "r", encoding="utf-8") as f:
        items = json.load(f)
    items = sorted(items, key=lambda it: int(it["object_id"]))
    base = Path(results_json_path).parent
    images = [Image.open(str(base / it["filename"])).convert("RGBA") for it in items]
    labels = [str(it.get("label", f"id_{it['object_id']}")) for it in items]
    return contact_sheet(images, labels, thumb_size, cols, label_height, font_size)


# ---------------------------------------------------------------------------- candidates grid
def candidates_grid(images: Sequence[Image.Image]) -> Optional[Image.Image]:
    """2 x 2 grid of up to four RGBA drafts, each LANCZOS-resized to the first one's size, on white."""
    if not images:
        return None
    ref_w, ref_h = images[0].size
    grid = Image.new("RGBA", (ref_w * 2, ref_h * 2), (255, 255, 255, 255))
    positions = [(0, 0), (ref_w, 0), (0, ref_h), (ref_w, ref_h)]
    objects = {i: im for i, im in enumerate(images)}
    placements = [{"object_id": i, "box": [x, y, x + ref_w, y + ref_h]} for i, (_, (x, y)) in enumerate(zip(images, positions))]
    return _compositor.composite(grid, objects, placements)


def compose_candidates_grid(image_paths: List[Path], out_path: Path) -> None:
    """Drop-in for ``_compose_candidates_grid`` (macro_placement_test.py:1332-1345)."""
    imgs = [Image.open(p).convert("RGBA") for p in image_paths if Path(p).exists()]
    grid = candidates_grid(imgs)
    if grid is not None:
        grid.save(out_path)


# ---------------------------------------------------------------------------- agentic compositor node
def composite_native_size(background_img: Image.Image, object_images, placements) -> Image.Image:
    """The raster loop of the agentic compositor node (agentic/nodes/compositor.py:36-43): every overlay must
    already have its placement's size -- ``ValueError("Placement size mismatch; scaling objects is not
    permitted")`` otherwise -- and is alpha-composited at (x, y) in list order.  ``placements`` are
    ``{object_id, box}`` records (the node builds exactly these, :22-34) or objects with
    ``object_id / x / y / width / height`` attributes.  One fused launch instead of one pass per overlay."""
    recs = []
    for p in placements:
        if isinstance(p, dict):
            oid, (x1, y1, x2, y2) = p["object_id"], p["box"]
            x, y, w, h = x1, y1, x2 - x1, y2 - y1
        else:
            oid, x, y, w, h = p.object_id, p.x, p.y, p.width, p.height
        if object_images[oid].size != (w, h):
            raise ValueError("Placement size mismatch; scaling objects is not permitted")
        recs.append({"object_id": oid, "box": [x, y, x + w, y + h]})
    return _compositor.composite(background_img, object_images, recs)
