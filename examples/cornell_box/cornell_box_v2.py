"""Headless run of the reference's examples/cornell_box/cornell_box_v2.py (x10 world, rounded boxes, plain marcher)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _common import run
from raytracingpbr_b200 import scenes

if __name__ == "__main__":
    run(scenes.cornell_box_v2, (512, 512), 64, "cornell_box_v2.png")                 # cornell_box_v2.py:7
