"""Extra pin: the oracle vs the live Pillow / NumPy installed in the image (the third-party
code the reference path delegates to; not /root/reference).  Seeded random sweep, CPU only."""
import numpy as np
import pytest

import oracle

PIL = pytest.importorskip("PIL")
from PIL import Image  # noqa: E402


def _soft(rng, h, w):
    a = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    r = np.hypot((xx - w / 2) / (w / 2 + 1e-9), (yy - h / 2) / (h / 2 + 1e-9))
    a[..., 3] = np.clip((1.0 - r) * 255 / 0.15, 0, 255).astype(np.uint8)
    return a


def test_resize_sweep_vs_pillow():
    rng = np.random.default_rng(7)
    for _ in range(60):
        sw, sh = (int(v) for v in rng.integers(1, 180, 2))
        w, h = (int(v) for v in rng.integers(1, 180, 2))
        src = _soft(rng, sh, sw) if rng.random() < 0.5 else rng.integers(0, 256, (sh, sw, 4), dtype=np.uint8)
        exp = np.array(Image.fromarray(src, "RGBA").resize((w, h), Image.LANCZOS))
        assert np.array_equal(oracle.resize_rgba_lanczos(src, (w, h)), exp), (sw, sh, w, h)


def test_large_resize_vs_pillow():
    rng = np.random.default_rng(8)
    src = _soft(rng, 700, 900)
    for size in [(611, 503), (450, 350), (900, 351), (1200, 800)]:
        exp = np.array(Image.fromarray(src, "RGBA").resize(size, Image.LANCZOS))
        assert np.array_equal(oracle.resize_rgba_lanczos(src, size), exp), size


def test_composite_sweep_vs_pillow():
    rng = np.random.default_rng(9)
    for _ in range(10):
        W, H = (int(v) for v in rng.integers(20, 300, 2))
        bg = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
        if rng.random() < 0.6:
            bg[..., 3] = 255
        objs = {i: _soft(rng, int(rng.integers(1, 120)), int(rng.integers(1, 120))) for i in range(4)}
        pl = []
        for _ in range(6):
            oid = int(rng.integers(0, 4))
            x1, y1 = int(rng.integers(-40, W)), int(rng.integers(-40, H))
            if rng.random() < 0.3:
                w, h = objs[oid].shape[1], objs[oid].shape[0]
            else:
                w, h = int(rng.integers(1, 200)), int(rng.integers(1, 200))
            pl.append({"object_id": oid, "box": [x1, y1, x1 + w, y1 + h]})
        canvas = Image.fromarray(bg, "RGBA")
        for p in pl:
            x1, y1, x2, y2 = p["box"]
            r = Image.fromarray(objs[p["object_id"]], "RGBA").resize((max(1, x2 - x1), max(1, y2 - y1)), Image.LANCZOS)
            canvas.alpha_composite(r, dest=(x1, y1)) if x1 >= 0 and y1 >= 0 else _neg_dest(canvas, r, x1, y1)
        assert np.array_equal(oracle.composite(bg, objs, pl), np.array(canvas))


def _neg_dest(canvas, overlay, x, y):
    """Image.alpha_composite rejects negative dest only via its source-box check in some
    versions; emulate the clip with a crop so the sweep can include negative destinations."""
    sx, sy = max(0, -x), max(0, -y)
    if sx >= overlay.size[0] or sy >= overlay.size[1]:
        return
    canvas.alpha_composite(overlay.crop((sx, sy, overlay.size[0], overlay.size[1])), dest=(max(0, x), max(0, y)))


def test_median_vs_numpy():
    rng = np.random.default_rng(10)
    for n in (1, 2, 7, 1000, 1001):
        a = rng.integers(0, 256, (1, n, 4), dtype=np.uint8)
        a[..., 3] = np.where(rng.random((1, n)) < 0.5, 0, 200)
        a[0, 0, 3] = 9
        m = a[..., 3] > 0
        exp = tuple(int(x) for x in np.median(a[..., :3][m], axis=0).tolist())
        assert oracle.masked_median_rgb(a) == exp
