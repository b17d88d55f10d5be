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
