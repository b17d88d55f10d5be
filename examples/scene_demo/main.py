"""Headless run of the reference's examples/scene_demo/main.py (gradient sky)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _common import run
from raytracingpbr_b200 import scenes

if __name__ == "__main__":
    run(scenes.scene_demo, (480, 270), 128, "scene_demo.png")                        # main.py:9
