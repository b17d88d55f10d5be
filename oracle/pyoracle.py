"""ctypes loader for the CPU oracle (oracle/oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

FAMILY_A, FAMILY_B, FAMILY_C = 0, 1, 2
MARCH_PLAIN, MARCH_ENHANCED, MARCH_SRC = 0, 1, 2
SKY_BLACK, SKY_ENVMAP, SKY_GRADIENT = 0, 1, 2
SHAPE_NONE, SHAPE_SPHERE, SHAPE_BOX, SHAPE_CYLINDER, SHAPE_CONE, SHAPE_PLANE, SHAPE_BUNNY = range(7)


class OrcObject(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("position", C.c_float * 3),
        ("rotation", C.c_float * 3),
        ("scale", C.c_float * 3),
        ("albedo", C.c_float * 3),
        ("emission", C.c_float * 3),
        ("roughness", C.c_float),
        ("metallic", C.c_float),
        ("transmission", C.c_float),
        ("ior", C.c_float),
    ]


class OrcConfig(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32),
        ("family", C.c_int32),
        ("max_bounces", C.c_int32),
        ("max_steps", C.c_int32),
        ("marcher", C.c_int32),
        ("t_start", C.c_float), ("hit_eps", C.c_float), ("t_far", C.c_float),
        ("relax_w0", C.c_float), ("relax_guard", C.c_int32), ("relax_reset", C.c_int32),
        ("relax_w_reset", C.c_float),
        ("normal_h", C.c_float), ("box_round", C.c_float), ("light_quality", C.c_float),
        ("bsdf", C.c_int32), ("f0_variant", C.c_int32),
        ("visibility_min", C.c_float), ("visibility_max", C.c_float),
        ("sky", C.c_int32), ("sky_scale", C.c_float),
        ("lookfrom", C.c_float * 3), ("lookat", C.c_float * 3), ("vup", C.c_float * 3),
        ("vfov", C.c_float), ("aspect", C.c_float), ("aperture", C.c_float), ("focus", C.c_float),
        ("seed", C.c_uint32),
        ("frame", C.c_int32),
        ("min_dis", C.c_float), ("pixel_radius", C.c_float), ("quality_per_sample", C.c_float),
        ("black_background", C.c_int32),
        ("nearest_seed", C.c_int32), ("normal_mode", C.c_int32), ("samples_per_pixel", C.c_int32),
        ("adaptive_sampling", C.c_int32), ("noise_threshold", C.c_float),
        ("inner_spp", C.c_int32), ("primary_miss", C.c_int32), ("bunny_bob", C.c_int32),
    ]


class OrcCounters(C.Structure):
    _fields_ = [("scene_evals", C.c_uint64), ("rays", C.c_uint64), ("normals", C.c_uint64), ("samples", C.c_uint64)]


def build(force: bool = False) -> str:
    """Compile oracle/liboracle.so with the committed Makefile (gcc, OpenMP)."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle.so"])
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        f32p = C.POINTER(C.c_float)
        L.orc_pathtrace.restype = C.c_int
        L.orc_pathtrace.argtypes = [C.POINTER(OrcConfig), C.POINTER(OrcObject), C.c_int, f32p, C.c_int, C.c_uint32,
                                    C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(OrcCounters)]
        L.orc_pathtrace_ex.restype = C.c_int
        L.orc_pathtrace_ex.argtypes = [C.POINTER(OrcConfig), C.POINTER(OrcObject), C.c_int, f32p, f32p, f32p, C.c_int, C.c_int,
                                       C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(OrcCounters)]
        L.orc_pathtrace_adaptive.restype = C.c_int
        L.orc_pathtrace_adaptive.argtypes = [C.POINTER(OrcConfig), C.POINTER(OrcObject), C.c_int, f32p, f32p, f32p, f32p, C.c_int,
                                             C.c_int, C.c_int, C.c_uint32, C.c_int]
        L.orc_post_process_src.restype = None
        L.orc_post_process_src.argtypes = [C.c_int, f32p, f32p, f32p, f32p, C.c_float, C.c_double, C.c_int]
        L.orc_pathtrace_cols.restype = C.c_int
        L.orc_pathtrace_cols.argtypes = [C.POINTER(OrcConfig), C.POINTER(OrcObject), C.c_int, f32p, C.c_int, C.c_uint32,
                                         C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int]
        L.orc_signed_distance_g.restype = C.c_float
        L.orc_signed_distance_g.argtypes = [C.POINTER(OrcConfig), C.POINTER(OrcObject), C.c_int, C.c_int, f32p]
        L.orc_nearest_g.restype = C.c_int
        L.orc_nearest_g.argtypes = [C.POINTER(OrcConfig), C.POINTER(OrcObject), C.c_int, f32p, f32p]
        L.orc_calc_normal_g.restype = None
        L.orc_calc_normal_g.argtypes = [C.POINTER(OrcConfig), C.POINTER(OrcObject), C.c_int, C.c_int, f32p, f32p]
        L.orc_raycast_g.restype = None
        L.orc_raycast_g.argtypes = [C.POINTER(OrcConfig), C.POINTER(OrcObject), C.c_int, f32p, f32p, f32p]
        L.orc_sd_bunny.restype = C.c_float
        L.orc_sd_bunny.argtypes = [f32p]
        L.orc_sky_envmap.restype = None
        L.orc_sky_envmap.argtypes = [f32p, C.c_int, C.c_int, f32p, f32p]
        L.orc_sd_box.restype = C.c_float
        L.orc_sd_box.argtypes = [f32p, f32p, C.c_float]
        L.orc_signed_distance.restype = C.c_float
        L.orc_signed_distance.argtypes = [C.POINTER(OrcConfig), C.POINTER(OrcObject), C.c_int, C.c_int, f32p, C.c_int]
        L.orc_nearest.restype = C.c_int
        L.orc_nearest.argtypes = [C.POINTER(OrcConfig), C.POINTER(OrcObject), C.c_int, f32p, f32p]
        L.orc_calc_normal.restype = None
        L.orc_calc_normal.argtypes = [C.POINTER(OrcConfig), C.POINTER(OrcObject), C.c_int, C.c_int, f32p, f32p]
        L.orc_raycast.restype = None
        L.orc_raycast.argtypes = [C.POINTER(OrcConfig), C.POINTER(OrcObject), C.c_int, f32p, f32p, f32p]
        L.orc_hemispheric_sampling.restype = None
        L.orc_hemispheric_sampling.argtypes = [f32p, C.c_float, C.c_float, f32p]
        L.orc_camera_ray.restype = None
        L.orc_camera_ray.argtypes = [C.POINTER(OrcConfig), C.c_int, C.c_int, C.c_float, C.c_float, f32p, f32p]
        L.orc_rr_prob.restype = C.c_float
        L.orc_rr_prob.argtypes = [C.POINTER(OrcConfig), C.c_int]
        L.orc_sincosf.restype = None
        L.orc_sincosf.argtypes = [C.c_float, f32p, f32p]
        L.orc_atan2f.restype = C.c_float
        L.orc_atan2f.argtypes = [C.c_float, C.c_float]
        L.orc_asinf.restype = C.c_float
        L.orc_asinf.argtypes = [C.c_float]
        L.orc_angle_deg.restype = None
        L.orc_angle_deg.argtypes = [f32p, f32p]
        L.orc_philox4x32_10.restype = None
        L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.orc_random.restype = C.c_float
        L.orc_random.argtypes = [C.c_uint32] * 4
        L.orc_draw4.restype = None
        L.orc_draw4.argtypes = [C.c_uint32] * 5 + [f32p]
        assert L.orc_sizeof_config() == C.sizeof(OrcConfig), "OrcConfig layout mismatch"
        assert L.orc_sizeof_object() == C.sizeof(OrcObject), "OrcObject layout mismatch"
        _lib = L
    return _lib


def _f32p(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def vec(x, y=None, z=None):
    if y is None:
        return (C.c_float * 3)(x, x, x)
    return (C.c_float * 3)(x, y, z)


def make_object(type_, position, rotation, scale, albedo, emission=(1, 1, 1), roughness=1.0, metallic=0.0,
                transmission=0.0, ior=1.0) -> OrcObject:
    o = OrcObject()
    o.type = int(type_)
    o.position[:] = [np.float32(v) for v in position]
    o.rotation[:] = [np.float32(v) for v in rotation]
    o.scale[:] = [np.float32(v) for v in scale]
    o.albedo[:] = [np.float32(v) for v in albedo]
    o.emission[:] = [np.float32(v) for v in emission]
    o.roughness, o.metallic, o.transmission, o.ior = roughness, metallic, transmission, ior
    return o


def objects_array(objs):
    arr = (OrcObject * len(objs))()
    for i, o in enumerate(objs):
        arr[i] = o
    return arr


def cornell_shortest_objects():
    """Scene of examples/cornell_box/cornell_box_shortest.py:17-32."""
    f = np.float32
    w = [f(1) * f(0.4)] * 3
    B = SHAPE_BOX
    objs = [
        make_object(B, (0, 0, -1), (0, 0, 0), (1, 1, 0.2), w),
        make_object(B, (0, 1, 0), (90, 0, 0), (1, 1, 0.2), w),
        make_object(B, (0, -1, 0), (90, 0, 0), (1, 1, 0.2), w),
        make_object(B, (-1, 0, 0), (0, 90, 0), (1, 1, 0.2), (f(1) * f(0.5), 0, 0)),
        make_object(B, (1, 0, 0), (0, 90, 0), (1, 1, 0.2), (0, f(1) * f(0.5), 0)),
        make_object(B, (-0.275, -0.3, -0.2), (0, 112, 0), (0.25, 0.5, 0.25), w),
        make_object(B, (0.275, -0.55, 0.2), (0, -197, 0), (0.25, 0.25, 0.25), w),
        make_object(B, (0, 0.809, 0), (90, 0, 0), (0.2, 0.2, 0.01), (1, 1, 1), (100, 100, 100)),
    ]
    return objs


def cornell_shortest_config(width=512, height=512, max_bounces=3, seed=0) -> OrcConfig:
    """Constants of examples/cornell_box/cornell_box_shortest.py (family A)."""
    c = OrcConfig()
    c.width, c.height = width, height
    c.family = FAMILY_A
    c.max_bounces = max_bounces          # :83
    c.max_steps = 256                    # :66
    c.marcher = MARCH_PLAIN
    c.t_start, c.hit_eps, c.t_far = 0.0005, 0.00001, 2000.0   # :65,70,71
    c.relax_w0, c.relax_guard, c.relax_reset, c.relax_w_reset = 1.0, 0, 0, 1.0
    c.normal_h = float(np.float32(0.5773) * np.float32(0.005))  # :57
    c.box_round = 0.0
    c.light_quality = 128.0              # :84
    c.bsdf, c.f0_variant = 0, 0
    c.visibility_min, c.visibility_max = 0.000001, float("inf")  # :99
    c.sky, c.sky_scale = SKY_BLACK, 1.0
    c.lookfrom[:] = (0, 0, 3.5)          # :135
    c.lookat[:] = (0, 0, -1)
    c.vup[:] = (0, 1, 0)
    c.vfov, c.aspect, c.aperture, c.focus = 35.0, 1.0, 0.0, 1.0   # :111
    c.seed = seed
    c.frame = 0
    c.min_dis, c.pixel_radius, c.quality_per_sample, c.black_background = 0.0, 0.0, 0.8, 0
    c.nearest_seed, c.normal_mode, c.samples_per_pixel = 0, 0, 1
    c.adaptive_sampling, c.noise_threshold = 0, 1e-4
    return c


def pathtrace(cfg: OrcConfig, objs, spp: int, sample_base: int = 0, image_buffer: np.ndarray | None = None,
              i0: int = 0, i1: int | None = None, hoisted: bool = True, nthreads: int = 0, counters: bool = False,
              ray_buffer: np.ndarray | None = None, env: np.ndarray | None = None):
    """Accumulate `spp` samples per pixel (families A/B) or run `spp` launches of kernel
    pathtrace() (family C, state in `ray_buffer` (W,H,10) f32 with depth as raw int32 bits);
    returns the (W,H,4) f32 buffer (and counters).  `env` is the processed (w,h,3) f32 table."""
    L = lib()
    W, H = cfg.width, cfg.height
    if image_buffer is None:
        image_buffer = np.zeros((W, H, 4), dtype=np.float32)
    assert image_buffer.shape == (W, H, 4) and image_buffer.dtype == np.float32 and image_buffer.flags.c_contiguous
    if cfg.family == FAMILY_C:
        assert ray_buffer is not None and ray_buffer.shape == (W, H, 10) and ray_buffer.dtype == np.float32
    if env is not None:
        assert env.dtype == np.float32 and env.flags.c_contiguous and env.ndim == 3 and env.shape[2] == 3
    arr = objs if isinstance(objs, C.Array) else objects_array(objs)
    cnt = OrcCounters()
    rc = L.orc_pathtrace_ex(C.byref(cfg), arr, len(arr), _f32p(image_buffer),
                            _f32p(ray_buffer) if ray_buffer is not None else None,
                            _f32p(env) if env is not None else None, env.shape[0] if env is not None else 0,
                            env.shape[1] if env is not None else 0, spp, sample_base, i0,
                            W if i1 is None else i1, int(hoisted), nthreads, C.byref(cnt) if counters else None)
    if rc != 0:
        raise RuntimeError(f"orc_pathtrace_ex failed: {rc}")
    if counters:
        return image_buffer, {k: getattr(cnt, k) for k, _ in OrcCounters._fields_}
    return image_buffer


def pathtrace_columns(cfg: OrcConfig, objs, spp: int, columns, image_buffer: np.ndarray, hoisted: bool = True,
                      nthreads: int = 0, sample_base: int = 0) -> np.ndarray:
    """Families A/B on an explicit list of columns, one parallel region (bench.py's bounded CPU sample)."""
    arr = objs if isinstance(objs, C.Array) else objects_array(objs)
    cols = (C.c_int * len(columns))(*[int(c) for c in columns])
    rc = lib().orc_pathtrace_cols(C.byref(cfg), arr, len(arr), _f32p(image_buffer), spp, sample_base, cols, len(columns),
                                  int(hoisted), nthreads)
    if rc != 0:
        raise RuntimeError(f"orc_pathtrace_cols failed: {rc}")
    return image_buffer


def pathtrace_adaptive(cfg: OrcConfig, objs, launches: int, image_buffer, ray_buffer, diff_pixels, env, sample_base: int = 0):
    """Family C with ADAPTIVE_SAMPLING: `launches` x kernel pathtrace() reading the diff_pixels field."""
    arr = objs if isinstance(objs, C.Array) else objects_array(objs)
    rc = lib().orc_pathtrace_adaptive(C.byref(cfg), arr, len(arr), _f32p(image_buffer), _f32p(ray_buffer), _f32p(diff_pixels),
                                      _f32p(env) if env is not None else None, env.shape[0] if env is not None else 0,
                                      env.shape[1] if env is not None else 0, launches, sample_base, 0)
    if rc != 0:
        raise RuntimeError(f"orc_pathtrace_adaptive failed: {rc}")


def denoise(pixels_in: np.ndarray, out_prev: np.ndarray, threshold: float) -> np.ndarray:
    """One deterministic pass of kernel denoise() (examples/denoise/denoise_test_1.py:86-118); returns the new output."""
    W, H, _ = pixels_in.shape
    pixels_in = np.ascontiguousarray(pixels_in, dtype=np.float32)
    out_prev = np.ascontiguousarray(out_prev, dtype=np.float32)
    out = np.empty_like(pixels_in)
    L = lib()
    L.orc_denoise.restype = None
    L.orc_denoise.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float]
    L.orc_denoise(W, H, _f32p(pixels_in), _f32p(out_prev), _f32p(out), threshold)
    return out


def post_process_src(image_buffer, image_pixels, diff_buffer, diff_pixels, exposure: float, gamma: float, adaptive: bool):
    """kernel post_process() of src/postprocessor.py:24-43, in place on image_pixels / diff_*."""
    n = image_buffer.shape[0] * image_buffer.shape[1]
    lib().orc_post_process_src(n, _f32p(image_buffer), _f32p(image_pixels), _f32p(diff_buffer), _f32p(diff_pixels),
                               exposure, gamma, int(adaptive))
