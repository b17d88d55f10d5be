"""Shared helpers for the test-suite: product structs -> oracle structs, host-check loader."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import pyoracle as po  # noqa: E402  (test infrastructure)
from raytracingpbr_b200 import _native as N  # noqa: E402
from raytracingpbr_b200.dataclass import Camera, SDFObject  # noqa: E402


def to_oracle(cfg: N.RtpbrConfig, camera: Camera, objects):
    """Translate the product's (config, camera, objects) into the oracle's own structs."""
    oc = po.OrcConfig()
    for name, _ in N.RtpbrConfig._fields_:
        if name in ("kernel", "count_work"):
            continue
        setattr(oc, name, getattr(cfg, name))
    nc = camera.to_native() if isinstance(camera, Camera) else camera
    oc.lookfrom[:] = list(nc.lookfrom)
    oc.lookat[:] = list(nc.lookat)
    oc.vup[:] = list(nc.vup)
    oc.vfov, oc.aspect, oc.aperture, oc.focus = nc.vfov, nc.aspect, nc.aperture, nc.focus
    oc.frame = 0
    oobjs = []
    for o in objects:
        n = o.to_native() if isinstance(o, SDFObject) else o
        oo = po.OrcObject()
        for name, _ in po.OrcObject._fields_:
            v = getattr(n, name)
            if hasattr(v, "__len__"):
                getattr(oo, name)[:] = list(v)
            else:
                setattr(oo, name, v)
        oobjs.append(oo)
    return oc, po.objects_array(oobjs)


_HC = None
_HC_DIR = os.path.join(ROOT, "tests", "native")


def hostcheck() -> C.CDLL:
    """Build (if stale) and load tests/native/libhostcheck.so: the product's
    __host__ __device__ integrator code compiled for the CPU (test harness only)."""
    global _HC
    if _HC is not None:
        return _HC
    so = os.path.join(_HC_DIR, "libhostcheck.so")
    src = os.path.join(_HC_DIR, "hostcheck.cu")
    csrc = os.path.join(ROOT, "raytracingpbr_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".h"))]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call([
            "nvcc", "-O2", "-std=c++17", "-shared", "-Wno-deprecated-gpu-targets", "-Xcompiler",
            "-fPIC,-ffp-contract=off,-fno-fast-math,-mfma,-fopenmp,-fvisibility=hidden", "-o", so, src, "-lgomp"],
            stderr=subprocess.DEVNULL)
    L = C.CDLL(so)
    L.hostcheck_pathtrace.restype = C.c_int
    L.hostcheck_pathtrace.argtypes = [C.POINTER(N.RtpbrConfig), C.POINTER(N.RtpbrCamera), C.POINTER(N.RtpbrObject), C.c_int,
                                      C.POINTER(C.c_float), C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_int]
    L.hostcheck_pathtrace_ex.restype = C.c_int
    f32p = C.POINTER(C.c_float)
    L.hostcheck_pathtrace_ex.argtypes = [C.POINTER(N.RtpbrConfig), C.POINTER(N.RtpbrCamera), C.POINTER(N.RtpbrObject), C.c_int,
                                         f32p, f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                         f32p]
    L.hostcheck_sd_bunny.restype = C.c_float
    L.hostcheck_sd_bunny.argtypes = [f32p]
    L.hostcheck_sincos.restype = None
    L.hostcheck_sincos.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.hostcheck_atan2.restype = C.c_float
    L.hostcheck_atan2.argtypes = [C.c_float, C.c_float]
    L.hostcheck_asin.restype = C.c_float
    L.hostcheck_asin.argtypes = [C.c_float]
    L.hostcheck_euler.restype = None
    L.hostcheck_euler.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float)]
    _HC = L
    return L


def hostcheck_pathtrace(cfg, camera, objects, spp, sample_base=0, image=None, rank=0, nranks=1, band=32,
                        ray_buffer=None, env=None, frame=0, diff_pixels=None):
    L = hostcheck()
    nobjs = [o.to_native() if isinstance(o, SDFObject) else o for o in objects]
    arr = (N.RtpbrObject * len(nobjs))(*nobjs)
    cam = camera.to_native() if isinstance(camera, Camera) else camera
    if image is None:
        image = np.zeros((cfg.width, cfg.height, 4), dtype=np.float32)
    f32p = C.POINTER(C.c_float)
    rc = L.hostcheck_pathtrace_ex(C.byref(cfg), C.byref(cam), arr, len(nobjs), image.ctypes.data_as(f32p),
                                  ray_buffer.ctypes.data_as(f32p) if ray_buffer is not None else None,
                                  env.ctypes.data_as(f32p) if env is not None else None,
                                  env.shape[0] if env is not None else 0, env.shape[1] if env is not None else 0,
                                  frame, spp, sample_base, rank, nranks, band,
                                  diff_pixels.ctypes.data_as(f32p) if diff_pixels is not None else None)
    assert rc == 0, rc
    return image


def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


# ------------------------------------------------------------------------------ golden fixtures
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def env_table(u8: np.ndarray, exposure: float, gamma: float) -> np.ndarray:
    """Independent restatement (test side) of Image.__init__/process (src/ibl.py:12-23,
    tokyo_ibl.py:40-51): (u8 / 255 * exposure) ** gamma, pow in binary64 rounded once."""
    x = (u8.astype(np.float32) / np.float32(255)) * np.float32(exposure)
    return np.power(x.astype(np.float64), float(np.float32(gamma))).astype(np.float32)


def golden_case(name: str):
    """(fixture, RtpbrConfig, objects, camera, tonemap, processed env table or None) of a golden fixture."""
    from raytracingpbr_b200 import scenes
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    W, H, seed = int(g["width"]), int(g["height"]), int(g["seed"])
    env = None
    if name.startswith("shortest"):
        cfg, objs, cam, tm = scenes.cornell_box_shortest(W, H, max_bounces=int(g["bounces"]), seed=seed)
    elif name == "cornell_box":
        cfg, objs, cam, tm = scenes.cornell_box(W, H, max_bounces=int(g["bounces"]), seed=seed)
    elif name == "cornell_v2":
        cfg, objs, cam, tm = scenes.cornell_box_v2(W, H, max_bounces=int(g["bounces"]), seed=seed)
    elif name == "cornell_v3":
        cfg, objs, cam, tm = scenes.cornell_box_v3(W, H, max_bounces=int(g["bounces"]), seed=seed)
    elif name == "tokyo_ibl":
        cfg, objs, cam, tm = scenes.tokyo_ibl(W, H, seed=seed)
        env = env_table(g["env_u8"], 1.8, 2.2)                      # tokyo_ibl.py:60
    elif name == "scene_demo":
        cfg, objs, cam, tm = scenes.scene_demo(W, H, seed=seed)
    elif name == "bunny_glass":
        cfg, objs, cam, tm = scenes.bunny_glass(W, H, max_bounces=int(g["bounces"]), seed=seed, frame=int(g["frame"]))
        env = env_table(g["env_u8"], 1.8, 2.2)                      # bunny_sdf_glass.py:279-280, applied per texel
    elif name in ("bunny_sdf_v2", "bunny_sdf"):
        preset = scenes.bunny_sdf_v2 if name == "bunny_sdf_v2" else scenes.bunny_sdf
        cfg, objs, cam, tm = preset(W, H, max_bounces=int(g["bounces"]), seed=seed, frame=int(g["frame"]), inner_spp=int(g["inner_spp"]))
        env = env_table(g["env_u8"], 1.8, 2.2)                      # bunny_sdf_v2.py:279-280
    elif name in ("src_scene", "src_adaptive"):
        cfg, objs, cam, tm = scenes.src_scene(W, H, seed=seed)
        env = env_table(g["env_u8"], 1.4, 2.2)                      # src/ibl.py:33
        if name == "src_adaptive":
            cfg.adaptive_sampling, cfg.noise_threshold = 1, float(g["noise_threshold"])
    else:
        raise KeyError(name)
    if "lookfrom" in g:
        cam.lookfrom, cam.lookat = g["lookfrom"], g["lookat"]
    return g, cfg, objs, cam, tm, env
