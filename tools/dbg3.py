import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import golden_io as G
from PIL import Image
from image_transformation_b200.compositor import composite
sys.path.insert(0, '.')
import oracle
bg, objs, pl, exp = G.case("c1_squarespace_1x1")
which = int(sys.argv[1])
pl = [pl[which]]
print("placement", pl, "cutout", objs[pl[0]["object_id"]].shape if pl[0]["object_id"] in objs else objs[str(pl[0]["object_id"])].shape)
out = np.array(composite(Image.fromarray(bg, "RGBA"), {k: Image.fromarray(v, "RGBA") for k, v in objs.items()}, pl))
exp = oracle.composite(bg, objs, pl)
print("mismatching pixels", int((out != exp).any(axis=2).sum()))
