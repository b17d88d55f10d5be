"""Headless run of the reference's examples/scene_demo/tokyo_ibl.py (7 objects, HDR environment)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _common import run
from raytracingpbr_b200 import scenes

if __name__ == "__main__":
    run(scenes.tokyo_ibl, (1920, 1080), 128, "tokyo_ibl.png", env=("Tokyo_BigSight_3k.hdr", 1.8, 2.2))   # tokyo_ibl.py:59-60
