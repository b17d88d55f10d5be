"""Headless run of the reference's examples/bunny/bunny_sdf_v2.py, frame 0 (kernel render(): SAMPLE_PER_PIXEL samples per
launch in an in-kernel loop; every launch is one finished frame)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _common import run
from raytracingpbr_b200 import scenes

if __name__ == "__main__":
    run(scenes.bunny_sdf_v2, (960, 540), 1, "bunny_sdf_v2.png", env=("limpopo_golf_course_3k.hdr", 1.8, 2.2), frame=0)
