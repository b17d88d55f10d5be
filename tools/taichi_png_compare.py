"""Compare converged renders of cornell_box_shortest (512 x 512) with the region means of the reference's Taichi-made
picture others/cornell_box_taichi.png (tests/golden/taichi_png_regions.npz) for several bounce caps.
    python tools/taichi_png_compare.py [spp] [bounces ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from raytracingpbr_b200 import PathTracer, scenes

g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "taichi_png_regions.npz"))
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
for spec in sys.argv[2:] or ["3"]:
    parts = spec.split(":")                              # [preset:]bounces[:tonemap mode:exposure]
    if parts[0].isdigit():
        parts = ["cornell_box_shortest"] + parts
    preset, bounces = parts[0], int(parts[1])
    cfg, objs, cam, tm = getattr(scenes, preset)(512, 512, max_bounces=bounces, seed=1)
    if len(parts) >= 4:
        tm = dict(mode=int(parts[2]), exposure=float(parts[3]), gamma=2.2)
    with PathTracer(cfg, objs, cam, tm) as pt:
        pt.refresh()
        pt.pathtrace(spp)
        pt.post_process()
        pix = pt.image_pixels.to_numpy()
    img = np.floor(np.clip(pix, 0, 1).transpose(1, 0, 2)[::-1] * 255.0)      # the orientation / quantisation of imwrite
    means = img.reshape(8, 64, 8, 64, 3).mean(axis=(1, 3))
    d = means - g["region_means"]
    print(f"{spec}: global mean {img.mean(axis=(0, 1)).round(2)} (png {g['global_mean'].round(2)}), region diff max {np.abs(d).max():.2f} "
          f"mean {np.abs(d).mean():.2f} rms {np.sqrt((d ** 2).mean()):.2f}")
    if os.environ.get("SHOW"):
        np.set_printoptions(precision=1, suppress=True, linewidth=200)
        print(d.mean(axis=2))
