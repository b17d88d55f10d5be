"""Headless run of the reference's examples/bunny/bunny_sdf_glass.py, frame 0 (neural SDF, glass)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _common import run
from raytracingpbr_b200 import scenes

if __name__ == "__main__":
    # the reference multiplies the texel by 1.8 and raises it to 2.2 at every lookup (bunny_sdf_glass.py:279-280)
    run(scenes.bunny_glass, (1024, 1024), 256, "bunny_sdf_glass.png", env=("limpopo_golf_course_3k.hdr", 1.8, 2.2), frame=0)
