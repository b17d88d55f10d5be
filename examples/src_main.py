"""Headless equivalent of the reference's `ti index.py` / src/main.py loop: build_scene(), then
render(refreshing) frame after frame, then save image_pixels (the 'g' key of src/main.py:53-56)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _common import find_asset
from raytracingpbr_b200 import imwrite
from raytracingpbr_b200.src import config

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--out", default="src_main.png")
    a = ap.parse_args()
    config.SAMPLES_PER_FRAME = 16
    from raytracingpbr_b200.src import ibl
    from raytracingpbr_b200.src.fileds import image_buffer, image_pixels
    from raytracingpbr_b200.src.renderer import render
    from raytracingpbr_b200.src.scene import build_scene
    hdr = find_asset("Tokyo_BigSight_3k.hdr")
    if hdr:
        ibl.load(hdr)                     # src/ibl.py:32-33
    build_scene()                         # src/main.py:21
    for frame in range(a.frames // config.SAMPLES_PER_FRAME):
        render(refreshing=(frame == 0))   # src/main.py:62
    imwrite(image_pixels.to_numpy(), a.out)
    print("accumulated paths:", float(image_buffer.to_numpy()[..., 3].sum()), "->", a.out)
