"""Headless run of the reference's examples/cornell_box/cornell_box.py (PBR materials, plain marcher)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _common import run
from raytracingpbr_b200 import scenes

if __name__ == "__main__":
    run(scenes.cornell_box, (480, 480), 64, "cornell_box.png")                       # cornell_box.py:6
