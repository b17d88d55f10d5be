"""The only artefact in the reference that Taichi itself produced: others/cornell_box_taichi.png, the README's picture of
examples/cornell_box/cornell_box_shortest.py (README.md:3-5).  tests/golden/taichi_png_regions.npz holds the mean colour
of its 8 x 8 regions (tests/tools/gen_taichi_png_fixture.py), not the picture.

What the comparison shows (tools/taichi_png_compare.py, 8192 spp):
  * the RADIANCE of this implementation (512 x 512, 3 bounces, as the script ships) reproduces the picture when it is
    tone-mapped  exposure -> ACES -> gamma  (the order of cornell_box.py:374-377 / cornell_box_v2.py): region means agree
    to 3.9 / 255 on average, global mean (110.2, 107.2, 77.5) against the picture's (107.4, 108.3, 75.3);
  * with the order the shipped file has (gamma -> ACES, cornell_box_shortest.py:124-129) the converged image is darker,
    global mean (92.6, 90.6, 64.0), whatever the bounce cap: the picture predates that revision of the tone-map tail.
So the picture pins geometry, light transport and brightness scale of the path tracer -- which is what this test asserts
-- but not the tone-map order of the shipped file."""
import os

import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu


def region_means(pix):
    img = np.floor(np.clip(pix, 0.0, 1.0).transpose(1, 0, 2)[::-1] * 255.0)       # orientation / quantisation of imwrite
    return img.reshape(8, 64, 8, 64, 3).mean(axis=(1, 3)), img.mean(axis=(0, 1))


def test_converged_cornell_box_matches_the_taichi_picture():
    from raytracingpbr_b200 import PathTracer, scenes
    g = np.load(os.path.join(common.GOLDEN_DIR, "taichi_png_regions.npz"))
    cfg, objs, cam, tm = scenes.cornell_box_shortest(512, 512, max_bounces=3, seed=1)      # as shipped: shortest:6, :83
    with PathTracer(cfg, objs, cam, tm) as pt:
        pt.refresh()
        pt.pathtrace(4096)                                                                 # 1.07 G samples, < 1 s
        pt.ctx.post_process(1, 1.0, 2.2)                                                   # exposure -> ACES -> gamma
        means, glob = region_means(pt.image_pixels.to_numpy())
        pt.post_process()                                                                  # the shipped order, for the record
        _, glob_shipped = region_means(pt.image_pixels.to_numpy())
    d = means - g["region_means"]
    assert np.abs(d).mean() < 6.0, np.abs(d).mean()          # measured 3.9
    assert np.abs(d).max() < 32.0, np.abs(d).max()           # measured 27.2 (one channel of one region on the coloured walls' edge)
    assert np.abs(glob - g["global_mean"]).max() < 5.0       # measured 2.8
    # red wall on the left, green wall on the right, like the picture
    left, right = means[2:6, 0], means[2:6, 7]
    gl, gr = g["region_means"][2:6, 0], g["region_means"][2:6, 7]
    assert (left[:, 0] > 2 * left[:, 1]).all() and (gl[:, 0] > 2 * gl[:, 1]).all()
    assert (right[:, 1] > 2 * right[:, 0]).all() and (gr[:, 1] > 2 * gr[:, 0]).all()
    # and the shipped tone-map order is the darker one (documented divergence of the picture, not of the kernel)
    assert glob_shipped.mean() < glob.mean() - 8.0
