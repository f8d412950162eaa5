import sys, numpy as np
sys.path.insert(0, '.')
from PIL import Image
from image_transformation_b200.compositor import composite
bg = Image.new("RGBA", (10, 10), (255, 0, 0, 255))
obj = Image.new("RGBA", (2, 2), (0, 255, 0, 255))
out = composite(bg, {1: obj}, [{"object_id": 1, "box": [4, 4, 6, 6]}])
print("pixel", out.getpixel((4, 4)), out.getpixel((0, 0)))
