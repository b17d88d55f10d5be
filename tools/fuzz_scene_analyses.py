#!/usr/bin/env python
"""GPU fuzz of the code generator's scene analyses: random rooms (axis-aligned slabs with quarter-turn rotations, freely
rotated boxes inside, a small light), families A and B (plain marcher, rounded boxes), rendered by the scene-specialised
kernel with the fast region + scene bounds + regeneration batches + finish threshold forced on, and by the ahead-of-time
kernel: the accumulation buffers must be bit-identical.

    python tools/fuzz_scene_analyses.py [n_scenes] [first_seed]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from raytracingpbr_b200 import PathTracer, _native as N, scenes
from raytracingpbr_b200.dataclass import Material, SDFObject, Transform
from raytracingpbr_b200.tmath import vec3

for k, v in dict(RTPBR_JIT_FAST="1", RTPBR_JIT_BBOX="1", RTPBR_REGEN_MIN="16", RTPBR_REGEN_IDLE="8", RTPBR_FIN_MIN="4").items():
    os.environ.setdefault(k, v)
n_scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 8
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 100
quarter = [0.0, 90.0, 180.0, 270.0, -90.0]


def euler(rot):
    """Rz @ Ry @ Rx of radians(rot) in double precision: only used to find which world axis each local axis looks along."""
    x, y, z = np.radians(rot)
    cx, sx, cy, sy, cz, sz = np.cos(x), np.sin(x), np.cos(y), np.sin(y), np.cos(z), np.sin(z)
    A = np.array([[cz, sz, 0], [-sz, cz, 0], [0, 0, 1]])
    B = np.array([[cy, 0, -sy], [0, 1, 0], [sy, 0, cy]])
    Cm = np.array([[1, 0, 0], [0, cx, sx], [0, -sx, cx]])
    return A @ B @ Cm


bad = fast = 0
for seed in range(seed0, seed0 + n_scenes):
    rng = np.random.default_rng(seed)
    half = float(rng.uniform(0.6, 3.0))
    thick = float(rng.uniform(0.05, 0.3)) * half
    centre = rng.uniform(-0.5, 0.5, 3) * float(rng.choice([0.0, 1.0]))
    objs = []
    for axis in range(3):
        for sign in (-1.0, 1.0):
            if axis == 2 and sign > 0:
                continue                                    # open front
            pos = centre.copy()
            pos[axis] += sign * half
            scale = [half * 1.2] * 3
            scale[axis] = thick
            rot = [float(rng.choice(quarter)) if rng.random() < 0.5 else 0.0 for _ in range(3)]
            perm = np.abs(euler(rot)).argmax(axis=1)
            objs.append(SDFObject(type=scenes.SHAPE_BOX, transform=Transform(vec3(*pos), vec3(*rot), vec3(*[scale[perm[r]] for r in range(3)])),
                                  material=Material(vec3(*rng.uniform(0.3, 0.8, 3)), vec3(1), 1.0, 0.0, 0.0, 1.5)))
    for _ in range(int(rng.integers(1, 4))):
        objs.append(SDFObject(type=scenes.SHAPE_BOX,
                              transform=Transform(vec3(*(centre + rng.uniform(-0.5, 0.5, 3) * half)), vec3(*rng.uniform(-180, 180, 3)),
                                                  vec3(*rng.uniform(0.1, 0.35, 3) * half)),
                              material=Material(vec3(0.5), vec3(1), 1.0, 0.0, 0.0, 1.5)))
    objs.append(SDFObject(type=scenes.SHAPE_BOX, transform=Transform(vec3(*(centre + np.array([0, 0.75 * half, 0]))), vec3(0, 0, 0),
                                                                     vec3(0.25 * half, 0.01 * half, 0.25 * half)),
                          material=Material(vec3(1), vec3(60), 1.0, 0.0, 0.0, 1.0)))
    for family in ("A", "B"):
        preset = scenes.cornell_box_shortest if family == "A" else scenes.cornell_box_v2
        cfg, _, cam, tm = preset(160, 120, max_bounces=6, seed=seed)
        cam.lookfrom = vec3(float(centre[0]), float(centre[1]), float(centre[2] + 3.5 * half))
        cam.lookat = vec3(float(centre[0]), float(centre[1]), float(centre[2]))
        src = N.jit_source(cfg, [o.to_native() for o in objs])
        out = {}
        for jit in (True, False):
            with PathTracer(cfg, objs, cam, tm) as pt:
                pt.ctx.set_jit(jit)
                pt.refresh()
                pt.pathtrace(4)
                out[jit] = pt.image_buffer.to_numpy()
                active, msg = pt.ctx.jit_status()
                assert active == jit, msg
        same = np.array_equal(out[True], out[False])
        lit = float((out[True][..., :3].sum(-1) > 0).mean())
        has_fast = "#define RT_JIT_FAST 1" in src
        fast += has_fast
        bad += not same
        print(f"seed {seed} family {family}: {len(objs)} boxes, fast region {'yes' if has_fast else 'no'}, scene bounds "
              f"{'yes' if '#define RT_JIT_BBOX 1' in src else 'no'}, lit pixels {lit:.2f}, specialised == ahead-of-time: {same}", flush=True)
print(f"{2 * n_scenes} renders, {fast} with a fast region, {bad} mismatches")
sys.exit(1 if bad else 0)
