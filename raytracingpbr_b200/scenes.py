"""Scene + parameter presets reproducing the reference's example scripts.

Each preset returns ``(RtpbrConfig, [SDFObject], Camera, tonemap)`` with resolution, spp and
max-bounces left as run-time parameters (BASELINE.json's configs
use other values than the files).
"""
from __future__ import annotations

import numpy as np

from . import _native as N
from .dataclass import Camera, Material, SDFObject, Transform
from .tmath import vec3

SHAPE_NONE, SHAPE_SPHERE, SHAPE_BOX, SHAPE_CYLINDER, SHAPE_CONE, SHAPE_PLANE, SHAPE_BUNNY = range(7)


def cornell_box_shortest(width: int = 512, height: int = 512, max_bounces: int = 3, seed: int = 0,
                         kernel: int = N.KERNEL_PERSISTENT, count_work: bool = False):
    """examples/cornell_box/cornell_box_shortest.py (family A): 8 boxes, diffuse only."""
    one = vec3(1)

    def box(pos, rot, scale, albedo, emission=one):
        # Material(albedo, emission) -- shortest:11; Transform(position, rotation, scale) -- shortest:12
        return SDFObject(type=SHAPE_BOX, transform=Transform(vec3(*pos), vec3(*rot), vec3(*scale)),
                         material=Material(albedo, emission, 1.0, 0.0, 0.0, 1.0))

    objects = [                                                                           # shortest:17-32
        box((0, 0, -1), (0, 0, 0), (1, 1, 0.2), vec3(1, 1, 1) * 0.4),                     # wall 1
        box((0, 1, 0), (90, 0, 0), (1, 1, 0.2), vec3(1, 1, 1) * 0.4),                     # wall 2
        box((0, -1, 0), (90, 0, 0), (1, 1, 0.2), vec3(1, 1, 1) * 0.4),                    # wall 3
        box((-1, 0, 0), (0, 90, 0), (1, 1, 0.2), vec3(1, 0, 0) * 0.5),                    # wall 4
        box((1, 0, 0), (0, 90, 0), (1, 1, 0.2), vec3(0, 1, 0) * 0.5),                     # wall 5
        box((-0.275, -0.3, -0.2), (0, 112, 0), (0.25, 0.5, 0.25), vec3(1, 1, 1) * 0.4),   # taller box
        box((0.275, -0.55, 0.2), (0, -197, 0), (0.25, 0.25, 0.25), vec3(1, 1, 1) * 0.4),  # box
        box((0, 0.809, 0), (90, 0, 0), (0.2, 0.2, 0.01), vec3(1, 1, 1) * 1, vec3(100)),   # light
    ]
    c = N.RtpbrConfig()
    c.width, c.height = width, height
    c.family = N.FAMILY_A
    c.max_bounces = max_bounces                      # range(3), shortest:83
    c.max_steps = 256                                # shortest:66
    c.marcher = N.MARCH_PLAIN
    c.t_start, c.hit_eps, c.t_far = 0.0005, 0.00001, 2000.0          # shortest:65,70,71
    c.relax_w0, c.relax_guard, c.relax_reset, c.relax_w_reset = 1.0, 0, 0, 1.0
    c.normal_h = float(np.float32(0.5773) * np.float32(0.005))       # shortest:57
    c.box_round = 0.0                                # shortest:45
    c.light_quality = 128.0                          # shortest:84
    c.bsdf, c.f0_variant = 0, 0
    c.visibility_min, c.visibility_max = 0.000001, float("inf")      # shortest:99
    c.sky, c.sky_scale = N.SKY_BLACK, 1.0            # shortest:89
    c.seed = seed
    c.min_dis, c.pixel_radius, c.quality_per_sample, c.black_background = 0.0, 0.0, 0.8, 0
    c.kernel = kernel
    c.count_work = int(count_work)
    camera = Camera(vec3(0, 0, 3.5), vec3(0, 0, -1), vec3(0, 1, 0), 35.0, 1.0, 0.0, 1.0)   # shortest:111,135
    tonemap = dict(mode=0, exposure=1.0, gamma=2.2)  # shortest:124-129
    return c, objects, camera, tonemap
