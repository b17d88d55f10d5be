"""Fixture from the one artefact in the reference that Taichi itself produced: others/cornell_box_taichi.png, the
README's picture of examples/cornell_box/cornell_box_shortest.py (README.md:3-5; 512 x 512, as the script ships).

Writes tests/golden/taichi_png_regions.npz: the mean 8-bit colour of each of the 8 x 8 regions of 64 x 64 pixels (row 0 =
top of the picture), i.e. 64 x 3 numbers -- not the picture.  Run here (the reference is not on the GPU box):

    python tests/tools/gen_taichi_png_fixture.py
"""
import os

import numpy as np
from PIL import Image

SRC = "/root/reference/others/cornell_box_taichi.png"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "golden", "taichi_png_regions.npz")

a = np.asarray(Image.open(SRC).convert("RGB")).astype(np.float64)          # (512, 512, 3), row 0 = top
assert a.shape == (512, 512, 3)
means = a.reshape(8, 64, 8, 64, 3).mean(axis=(1, 3))                        # (8, 8, 3)
np.savez_compressed(OUT, region_means=means, size=np.int32(512), grid=np.int32(8), source=SRC,
                    global_mean=a.mean(axis=(0, 1)))
print(OUT, means.shape, a.mean(axis=(0, 1)))

# The README's picture of examples/bunny/bunny_sdf_glass.py (others/sdf_bunny_glass.jpg, 1920 x 1080, some animation frame):
# the camera of that script is fixed, so everything outside the bunny is the environment seen through the thin lens.
SRC2 = "/root/reference/others/sdf_bunny_glass.jpg"
OUT2 = os.path.join(os.path.dirname(OUT), "taichi_bunny_jpg_regions.npz")
b = np.asarray(Image.open(SRC2).convert("RGB")).astype(np.float64)          # (1080, 1920, 3)
assert b.shape == (1080, 1920, 3)
np.savez_compressed(OUT2, region_means=b.reshape(8, 135, 8, 240, 3).mean(axis=(1, 3)), size=np.array([1920, 1080], np.int32),
                    grid=np.int32(8), source=SRC2, global_mean=b.mean(axis=(0, 1)))
print(OUT2, b.mean(axis=(0, 1)))
