import sys, numpy as np
sys.path.insert(0, '.')
from PIL import Image
from image_transformation_b200.compositor import composite
import oracle
W, H, w, h, x, y = (int(v) for v in sys.argv[1:7])
rng = np.random.default_rng(0)
bg = rng.integers(0, 256, (H, W, 4), dtype=np.uint8); bg[..., 3] = 255
ob = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
pl = [{"object_id": 1, "box": [x, y, x + w, y + h]}]
out = np.array(composite(Image.fromarray(bg, "RGBA"), {1: Image.fromarray(ob, "RGBA")}, pl))
exp = oracle.composite(bg, {1: ob}, pl)
print(sys.argv[1:7], "mismatching pixels", int((out != exp).any(axis=2).sum()))
