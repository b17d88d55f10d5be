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
    c.nearest_seed, c.normal_mode, c.samples_per_pixel = 0, 0, 1
    c.adaptive_sampling, c.noise_threshold = 0, 1e-4
    c.kernel = kernel
    c.count_work = int(count_work)
    camera = Camera(vec3(0, 0, 3.5), vec3(0, 0, -1), vec3(0, 1, 0), 35.0, 1.0, 0.0, 1.0)   # shortest:111,135
    tonemap = dict(mode=0, exposure=1.0, gamma=2.2)  # shortest:124-129
    return c, objects, camera, tonemap


# ------------------------------------------------------------------------------ helpers
def _f32(x):
    return np.float32(x)


def _base_config(width, height, seed, kernel, count_work):
    c = N.RtpbrConfig()
    c.width, c.height = width, height
    c.seed = seed
    c.kernel = kernel
    c.count_work = int(count_work)
    c.relax_w0, c.relax_guard, c.relax_reset, c.relax_w_reset = 1.0, 0, 0, 1.0
    c.visibility_min, c.visibility_max = 0.000001, float("inf")
    c.sky, c.sky_scale = N.SKY_BLACK, 1.0
    c.min_dis, c.pixel_radius, c.quality_per_sample, c.black_background = 0.0, 0.0, 0.8, 0
    c.nearest_seed, c.normal_mode, c.samples_per_pixel = 0, 0, 1
    c.adaptive_sampling, c.noise_threshold = 0, 1e-4      # src/config.py:14,17
    c.box_round = 0.0
    c.light_quality = 128.0
    c.f0_variant = 0
    return c


def _pixel_radius(width, height, k):
    """k * min(SCREEN_PIXEL_SIZE), SCREEN_PIXEL_SIZE = 1.0 / vec2(resolution) in f32 (tokyo_ibl.py:14-15)."""
    return float(_f32(k) * min(_f32(1.0) / _f32(width), _f32(1.0) / _f32(height)))


def _obj(type_, pos, rot, scale, albedo, emission, roughness, metallic, transmission, ior):
    return SDFObject(type=type_, transform=Transform(vec3(*pos), vec3(*rot), vec3(*scale)),
                     material=Material(albedo, emission, roughness, metallic, transmission, ior))


def _cornell_pbr_objects(tall_box_yaw, world_scale=1.0):
    """WORLD_LIST of cornell_box.py:161-190 (v2/v3 multiply position and scale by 10 at evaluation
    time, cornell_box_v3/sdf.py:17-19; done here once, in f32)."""
    w4, one = vec3(1, 1, 1) * 0.4, vec3(1)
    k = _f32(world_scale)
    rows = [
        ((0, 0, -1), (0, 0, 0), (1, 1, 0.2), w4, one, 1.530),
        ((0, 1, 0), (90, 0, 0), (1, 1, 0.2), w4, one, 1.530),
        ((0, -1, 0), (90, 0, 0), (1, 1, 0.2), w4, one, 1.530),
        ((-1, 0, 0), (0, 90, 0), (1, 1, 0.2), vec3(1, 0, 0) * 0.5, one, 1.530),
        ((1, 0, 0), (0, 90, 0), (1, 1, 0.2), vec3(0, 1, 0) * 0.5, one, 1.530),
        ((-0.275, -0.3, -0.2), (0, tall_box_yaw, 0), (0.25, 0.5, 0.25), w4, one, 1.530),
        ((0.275, -0.55, 0.2), (0, -197, 0), (0.25, 0.25, 0.25), w4, one, 1.530),
        ((0, 0.809, 0), (90, 0, 0), (0.2, 0.2, 0.01), vec3(1, 1, 1), vec3(100), 1.0),
    ]
    objs = []
    for pos, rot, scale, albedo, emission, ior in rows:
        p = vec3(*pos) * k if world_scale != 1.0 else vec3(*pos)
        sc = vec3(*scale) * k if world_scale != 1.0 else vec3(*scale)
        objs.append(_obj(SHAPE_BOX, p, rot, sc, albedo, emission, 1.0, 0.0, 0.0, ior))
    return objs


def cornell_box(width: int = 480, height: int = 480, max_bounces: int = 128, seed: int = 0,
                kernel: int = N.KERNEL_PERSISTENT, count_work: bool = False):
    """examples/cornell_box/cornell_box.py (family B): PBR materials, plain sphere tracing."""
    c = _base_config(width, height, seed, kernel, count_work)
    c.family = N.FAMILY_B
    c.max_bounces = max_bounces                      # MAX_RAYTRACE, cornell_box.py:20
    c.max_steps = 512                                # MAX_RAYMARCH :19
    c.marcher = N.MARCH_PLAIN                        # :213-223
    c.t_start, c.hit_eps, c.t_far = 0.005, 0.0001, 2000.0            # MIN_DIS, PRECISION, MAX_DIS :14-16
    c.normal_h = 0.0001                              # e = vec2(1, -1) * PRECISION :207
    c.bsdf, c.f0_variant = 1, 0                      # :257-290, F0 *= 2.0*F0 :275
    objects = _cornell_pbr_objects(-253)             # :180
    camera = Camera(vec3(0, 0, 3), vec3(0, 0, -1), vec3(0, 1, 0), 43.6, width / height, 0.01, 4.0)   # :30-34, :384
    tonemap = dict(mode=1, exposure=0.6, gamma=2.2)  # :374-377
    return c, objects, camera, tonemap


def cornell_box_v2(width: int = 512, height: int = 512, max_bounces: int = 3, seed: int = 0,
                   kernel: int = N.KERNEL_PERSISTENT, count_work: bool = False):
    """examples/cornell_box/cornell_box_v2.py (family B): cornell_box.py at x10 world scale with rounded boxes."""
    c = _base_config(width, height, seed, kernel, count_work)
    c.family = N.FAMILY_B
    c.max_bounces = max_bounces                      # MAX_RAYTRACE = 3, cornell_box_v2.py:21
    c.max_steps = 512                                # :20
    c.marcher = N.MARCH_PLAIN                        # :187-196
    c.t_start, c.hit_eps, c.t_far = 0.05, 0.001, 2000.0              # MIN_DIS, PRECISION, MAX_DIS :15-17
    c.normal_h = 0.001                               # e = vec2(1, -1) * PRECISION :180
    c.box_round = 0.01                               # :130
    c.bsdf, c.f0_variant = 1, 0
    objects = _cornell_pbr_objects(-253, world_scale=10.0)            # :134-136, :156
    camera = Camera(vec3(0, 0, 35), vec3(0, 0, -10), vec3(0, 1, 0), 35.0, width / height, 0.01, 4.0)   # :27-31, :346
    tonemap = dict(mode=1, exposure=1.0, gamma=2.2)
    return c, objects, camera, tonemap


def cornell_box_v3(width: int = 512, height: int = 512, max_bounces: int = 3, seed: int = 0,
                   kernel: int = N.KERNEL_PERSISTENT, count_work: bool = False):
    """examples/cornell_box/cornell_box_v3/ (family B): world x10, rounded boxes, enhanced sphere tracing."""
    c = _base_config(width, height, seed, kernel, count_work)
    c.family = N.FAMILY_B
    c.max_bounces = max_bounces                      # config.py:15
    c.max_steps = 512                                # config.py:14
    c.marcher = N.MARCH_ENHANCED                     # pathtracer.py:52-78
    c.t_start, c.t_far = 0.05, 2000.0                # config.py:9-10
    c.hit_eps = _pixel_radius(width, height, 0.5)    # PIXEL_RADIUS, config.py:7
    c.relax_w0, c.relax_guard, c.relax_reset, c.relax_w_reset = 1.6, 1, 0, 1.0    # pathtracer.py:56,64-67
    c.normal_h = float(_f32(0.5773 * 0.005))         # NORMAL_PRECISION (Python double product), config.py:11
    c.box_round = 0.01                               # sdf.py:11
    c.bsdf, c.f0_variant = 1, 0                      # pbr.py:50-51
    objects = _cornell_pbr_objects(-253, world_scale=10.0)            # scene.py, sdf.py:17-19
    camera = Camera(vec3(0, 0, 35), vec3(0, 0, -10), vec3(0, 1, 0), 35.0, width / height, 0.01, 4.0)   # config.py:21-24, main.py:14
    tonemap = dict(mode=3, exposure=1.0, gamma=2.2)
    return c, objects, camera, tonemap


def _demo_objects(variant: str):
    """7-object scene: src/scene.py:11-33 ('src') or scene_demo/tokyo_ibl.py:101-123 ('tokyo'),
    sorted by type like the reference (stable sort)."""
    one = vec3(1)
    if variant == "src":
        rows = [
            (SHAPE_SPHERE, (0, -100.501, 0), (100,) * 3, vec3(1, 1, 1) * 0.6, one, 1.0, 1.0, 0, 1.100),
            (SHAPE_SPHERE, (0, 0, 0), (0.5,) * 3, vec3(1, 1, 1) * 0.9, vec3(1, 10, 1), 0, 1, 0, 1.000),
            (SHAPE_SPHERE, (1, -0.2, 0), (0.3,) * 3, vec3(0.2, 0.2, 1) * 0.9, one, 0.2, 1, 0, 1.100),
            (SHAPE_SPHERE, (0.0, -0.2, 2), (0.3,) * 3, vec3(1, 1, 1) * 0.9, one, 0, 0, 1, 1.500),
            (SHAPE_CYLINDER, (-1.0, -0.2, 0), (0.3,) * 3, vec3(1.0, 0.2, 0.2) * 0.9, one, 0, 0, 0, 1.460),
            (SHAPE_BOX, (0, 0, 5), (2, 1, 0.2), vec3(1, 1, 0.2) * 0.9, one, 0, 1, 0, 0.470),
            (SHAPE_BOX, (0, 0, -2), (2, 1, 0.2), vec3(1, 1, 1) * 0.9, one, 0, 1, 0, 2.950),
        ]
    else:
        rows = [
            (SHAPE_SPHERE, (0, -100.501, 0), (100,) * 3, vec3(1, 1, 1) * 0.6, one, 1, 1, 0, 1.635),
            (SHAPE_SPHERE, (0, 0, 0), (0.5,) * 3, vec3(1, 1, 1), vec3(0.1, 1, 0.1) * 10, 1, 0, 0, 1),
            (SHAPE_SPHERE, (1, -0.2, 0), (0.3,) * 3, vec3(0.2, 0.2, 1), one, 0.2, 1, 0, 1.100),
            (SHAPE_SPHERE, (0.0, -0.2, 2), (0.3,) * 3, vec3(1, 1, 1) * 0.9, one, 0, 0, 1, 1.5),
            (SHAPE_CYLINDER, (-1.0, -0.2, 0), (0.3,) * 3, vec3(1.0, 0.2, 0.2), one, 0, 0, 0, 1.460),
            (SHAPE_BOX, (0, 0, 5), (2, 1, 0.2), vec3(1, 1, 0.2) * 0.9, one, 0, 1, 0, 0.470),
            (SHAPE_BOX, (0, 0, -2), (2, 1, 0.2), vec3(1, 1, 1) * 0.9, one, 0, 1, 0, 2.950),
        ]
    objs = [_obj(t, pos, (0, 0, 0), sc, al, em, r, m, tr, ior) for t, pos, sc, al, em, r, m, tr, ior in rows]
    return sorted(objs, key=lambda o: o.type)


def tokyo_ibl(width: int = 2880, height: int = 1620, max_bounces: int = 512, seed: int = 0,
              kernel: int = N.KERNEL_PERSISTENT, count_work: bool = False):
    """examples/scene_demo/tokyo_ibl.py (family B): spheres / cylinder / rounded boxes, HDR environment.
    The environment table is set separately (PathTracer.set_envmap; ibl.load_envmap)."""
    c = _base_config(width, height, seed, kernel, count_work)
    c.family = N.FAMILY_B
    c.max_bounces = max_bounces                      # tokyo_ibl.py:23
    c.max_steps = 512                                # :22
    c.marcher = N.MARCH_ENHANCED                     # :246-265
    c.t_start, c.t_far = 0.005, 2000.0               # :17-18
    c.hit_eps = _pixel_radius(width, height, 0.5)    # PIXEL_RADIUS :15
    c.relax_w0, c.relax_guard, c.relax_reset, c.relax_w_reset = 1.6, 0, 1, 1.0    # :247, :255-256
    c.normal_h = float(_f32(0.5773) * _f32(0.005))   # vec2(1, -1) * 0.5773 * 0.005 :239
    c.box_round = 0.03                               # :193
    c.bsdf, c.f0_variant = 1, 1                      # F0 = 2.0*(eta-1)/(eta+1); F0 *= F0 :318
    c.nearest_seed = 1                               # :222
    c.sky = N.SKY_ENVMAP                             # :274-277
    objects = _demo_objects("tokyo")
    camera = Camera(vec3(0, -0.2, 4), vec3(0, -0.2, 3), vec3(0, 1, 0), 30.0, width / height, 0.01, 4.0)   # :31-35, :444
    tonemap = dict(mode=1, exposure=1.0, gamma=2.2)  # :434-439
    return c, objects, camera, tonemap


def scene_demo(width: int = 480, height: int = 270, max_bounces: int = 128, seed: int = 0,
               kernel: int = N.KERNEL_PERSISTENT, count_work: bool = False):
    """examples/scene_demo/main.py: tokyo_ibl's scene under a gradient sky (x1.8), guarded relaxation."""
    c, objects, camera, tonemap = tokyo_ibl(width, height, max_bounces, seed, kernel, count_work)
    c.relax_guard, c.relax_reset, c.relax_w_reset = 1, 0, 1.0         # scene_demo/main.py:233-234
    c.sky, c.sky_scale = N.SKY_GRADIENT, 1.8                          # :246-248, :322
    return c, objects, camera, tonemap


def bunny_glass(width: int = 1920, height: int = 1080, max_bounces: int = 512, seed: int = 0, frame: int = 0,
                kernel: int = N.KERNEL_PERSISTENT, count_work: bool = False):
    """examples/bunny/bunny_sdf_glass.py (family B): neural-SDF glass bunny under an HDR environment.
    The table passed to set_envmap must already hold pow(texel * 1.8, 2.2) (bunny_sdf_glass.py:277-281)."""
    c = _base_config(width, height, seed, kernel, count_work)
    c.family = N.FAMILY_B
    c.max_bounces = max_bounces                      # :25
    c.max_steps = 2048                               # :24
    c.marcher = N.MARCH_ENHANCED                     # :248-267
    c.t_start, c.t_far = 0.005, 2000.0               # :18-19
    c.hit_eps = _pixel_radius(width, height, 0.5)    # :16
    c.relax_w0, c.relax_guard, c.relax_reset, c.relax_w_reset = 0.5, 1, 0, 0.4    # :251, :257-258
    c.normal_h = 0.0001                              # PRECISION :241
    c.light_quality = 512.0                          # :33
    c.bsdf, c.f0_variant = 1, 0                      # :322
    c.sky = N.SKY_ENVMAP
    c.bunny_bob = 1                                  # p += vec3(0, 0, 0.1*sin(t)) :215
    objects = [_obj(SHAPE_BUNNY, (0, 0, 0), (-90, 0, 0), (1, 1, 1), vec3(1, 1, 1) * 0.9, vec3(1), 0, 0, 1, 1.500)]   # :221-225
    camera = Camera(vec3(0, 0, 4), vec3(0, 0, 3), vec3(0, 1, 0), 30.0, width / height, 0.03, 4.0)     # :34-37, :435
    tonemap = dict(mode=1, exposure=0.8, gamma=2.2, frame=int(frame))  # :423-432; `frame`: u_frame (:409), applied by PathTracer
    return c, objects, camera, tonemap


def bunny_sdf_v2(width: int = 3840, height: int = 2160, max_bounces: int = 128, seed: int = 0, frame: int = 0, inner_spp: int = 12,
                 kernel: int = N.KERNEL_PERSISTENT, count_work: bool = False):
    """examples/bunny/bunny_sdf_v2.py (family B): the metal bunny on a white background.  Kernel render() traces
    SAMPLE_PER_PIXEL samples per launch in an in-kernel loop that shares one ti.random stream per pixel and overwrites
    image_buffer (:416-431): `pathtrace(n)` = n such launches.  Table for set_envmap: pow(texel * 1.8, 2.2) (:279-280)."""
    c, _, _, _ = bunny_glass(width, height, max_bounces, seed, frame, kernel, count_work)
    c.max_steps = 512                                # MAX_RAYMARCH :24
    c.relax_w0, c.relax_guard, c.relax_reset, c.relax_w_reset = 1.6, 1, 0, 0.7     # :251, :257-258
    c.light_quality = 128.0                          # :33
    c.inner_spp = inner_spp                          # SAMPLE_PER_PIXEL :23, loop :419
    c.primary_miss = 1                               # white background for camera rays :355-358
    c.bunny_bob = 1                                  # :213-216
    objects = [_obj(SHAPE_BUNNY, (0, 0, 0), (-90, 0, 0), (1, 1, 1), vec3(1, 1, 1) * 0.9, vec3(1), 0.0, 1, 0, 2.950)]   # :221-225
    camera = Camera(vec3(0, 0, 4), vec3(0, 0, 3), vec3(0, 1, 0), 30.0, width / height, 0.01, 4.0)      # :34-37, :437
    tonemap = dict(mode=1, exposure=0.8, gamma=2.2, frame=int(frame))      # :426-429 (exposure -> ACES -> gamma; the file does not clamp)
    return c, objects, camera, tonemap


def bunny_sdf(width: int = 3840, height: int = 2160, max_bounces: int = 128, seed: int = 0, frame: int = 0, inner_spp: int = 4,
              kernel: int = N.KERNEL_PERSISTENT, count_work: bool = False):
    """examples/bunny/bunny_sdf.py: like bunny_sdf_v2.py with SAMPLE_PER_PIXEL = 4, rotation without the bob (:214), camera
    rays that miss are black (`ray.color *= sign(float(i))`, :352), Tokyo environment (x 1.8, ^2.2), camera at z = 5."""
    c, objects, camera, tonemap = bunny_sdf_v2(width, height, max_bounces, seed, frame, inner_spp, kernel, count_work)
    c.primary_miss = 2
    c.bunny_bob = 0
    camera = Camera(vec3(0, 0, 5), vec3(0, 0, 4), vec3(0, 1, 0), 30.0, width / height, 0.01, 4.0)      # :432
    tonemap = dict(mode=1, exposure=0.6, gamma=2.2, frame=int(frame))      # camera_exposure :34
    return c, objects, camera, tonemap


def src_scene(width: int = 768, height: int = 432, max_bounces: int = 512, seed: int = 0,
              kernel: int = N.KERNEL_PERSISTENT, count_work: bool = False):
    """src/ (family C): progressive one-bounce-per-launch integrator, src/config.py + src/scene.py."""
    c = _base_config(width, height, seed, kernel, count_work)
    c.family = N.FAMILY_C
    c.max_bounces = max_bounces                      # MAX_RAYTRACE, src/config.py:26
    c.max_steps = 512                                # MAX_RAYMARCH :25
    c.marcher = N.MARCH_SRC                          # src/scene.py:59-84
    pr = _pixel_radius(width, height, 1.0)           # PIXEL_RADIUS = 1.0 * SCREEN_PIXEL_SIZE.min() :20
    c.pixel_radius = pr
    c.min_dis = float(_f32(2.5) * _f32(pr))          # MIN_DIS :22
    c.t_start, c.hit_eps, c.t_far = 0.0, pr, 1000.0  # MAX_DIS :23
    c.relax_w0, c.relax_guard, c.relax_reset, c.relax_w_reset = 1.6, 1, 0, 1.0
    c.normal_h = float(_f32(0.5773 * 0.005))         # Python-scope product, src/sdf.py:80
    c.normal_mode = 1                                # src/sdf.py:77-87
    c.box_round = 0.03                               # src/sdf.py:34
    c.bsdf, c.f0_variant = 2, 1                      # src/pbr.py:22-62
    c.nearest_seed = 1                               # src/scene.py:46
    c.visibility_min, c.visibility_max = 1e-4, 1e4   # VISIBILITY :16
    c.quality_per_sample = 0.8                       # :11
    c.black_background = 0                           # :13
    c.samples_per_pixel = 1                          # :10
    c.sky = N.SKY_ENVMAP                             # src/ibl.py:36-40
    objects = _demo_objects("src")
    aspect = float((_f32(1.0) / _f32(height)) / (_f32(1.0) / _f32(width)))    # SCREEN_PIXEL_SIZE.y / .x, src/camera.py:125
    camera = Camera(vec3(0, -0.2, 4.0), vec3(0, -0.2, 3.0), vec3(0, 1, 0), 35.0, aspect, 0.01, 4.0)   # src/camera.py:126-129, main.py:17
    tonemap = dict(mode=2, exposure=1.0, gamma=2.2)  # src/postprocessor.py:24-38
    return c, objects, camera, tonemap
