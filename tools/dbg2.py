import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import golden_io as G
from PIL import Image
from image_transformation_b200.compositor import composite
name = sys.argv[1] if len(sys.argv) > 1 else "c1_squarespace_1x1"
bg, objs, pl, exp = G.case(name)
out = np.array(composite(Image.fromarray(bg, "RGBA"), {k: Image.fromarray(v, "RGBA") for k, v in objs.items()}, pl))
print(name, "mismatching pixels", int((out != exp).any(axis=2).sum()), "of", exp.shape[0] * exp.shape[1])
