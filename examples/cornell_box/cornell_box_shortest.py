"""Headless run of the scene / algorithm of the reference's examples/cornell_box/cornell_box_shortest.py."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _common import run
from raytracingpbr_b200 import scenes

if __name__ == "__main__":
    run(scenes.cornell_box_shortest, (512, 512), 64, "cornell_box_shortest.png")     # shortest:6
