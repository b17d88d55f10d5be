"""ctypes binding of librtpbr.so (include/rtpbr.h).  No torch, no Taichi, no Triton.

The library is built in-tree by ``raytracingpbr_b200/csrc/Makefile`` (see ``build()``).
There is no CPU fallback: if the shared library is missing, or no sm_100 GPU is present,
the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librtpbr.so")
CSRC = os.path.join(_HERE, "csrc")

# enums of include/rtpbr.h
OK, ERR_ARG, ERR_CUDA, ERR_STATE, ERR_NCCL, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5
SHAPE_NONE, SHAPE_SPHERE, SHAPE_BOX, SHAPE_CYLINDER, SHAPE_CONE, SHAPE_PLANE, SHAPE_BUNNY = range(7)
FAMILY_A, FAMILY_B, FAMILY_C = 0, 1, 2
MARCH_PLAIN, MARCH_ENHANCED, MARCH_SRC = 0, 1, 2
SKY_BLACK, SKY_ENVMAP, SKY_GRADIENT = 0, 1, 2
KERNEL_PERSISTENT, KERNEL_SIMPLE = 0, 1
BUF_IMAGE_BUFFER, BUF_IMAGE_PIXELS, BUF_RAY_BUFFER, BUF_DIFF_BUFFER, BUF_DIFF_PIXELS, BUF_DENOISE_PIXELS = 0, 1, 2, 3, 4, 5
CNT_NAMES = ("scene_evals", "rays", "normals", "samples", "march_iters", "march_active", "resolve_rounds", "launches",
             "resolved_slots", "mlp_evals")

EXPORTS = (
    "rtpbr_create", "rtpbr_destroy", "rtpbr_set_scene", "rtpbr_set_camera", "rtpbr_set_envmap", "rtpbr_set_frame",
    "rtpbr_set_sample_base", "rtpbr_set_shard", "rtpbr_refresh", "rtpbr_pathtrace", "rtpbr_post_process",
    "rtpbr_download", "rtpbr_upload", "rtpbr_sync", "rtpbr_alloc_host", "rtpbr_free_host", "rtpbr_flush_l2", "rtpbr_timer_start", "rtpbr_timer_stop", "rtpbr_kernel_time",
    "rtpbr_get_counters", "rtpbr_device_info", "rtpbr_nccl_unique_id", "rtpbr_nccl_init", "rtpbr_reduce_tiles",
    "rtpbr_device_ptr", "rtpbr_set_jit", "rtpbr_jit_status", "rtpbr_jit_generate", "rtpbr_jit_compile_check", "rtpbr_last_error", "rtpbr_version", "rtpbr_sizeof_config", "rtpbr_sizeof_object",
    "rtpbr_sizeof_camera",
    "rtpbr_multi_create", "rtpbr_multi_destroy", "rtpbr_multi_count", "rtpbr_multi_context", "rtpbr_multi_set_scene",
    "rtpbr_multi_set_camera", "rtpbr_multi_set_envmap", "rtpbr_multi_set_frame", "rtpbr_multi_set_sample_base",
    "rtpbr_multi_refresh", "rtpbr_multi_pathtrace", "rtpbr_multi_reduce", "rtpbr_multi_post_process",
    "rtpbr_multi_download", "rtpbr_multi_sync", "rtpbr_device_count", "rtpbr_denoise",
)


class RtpbrError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"librtpbr error {code}: {msg}")
        self.code = code


class RtpbrObject(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("position", C.c_float * 3), ("rotation", C.c_float * 3), ("scale", C.c_float * 3),
        ("albedo", C.c_float * 3), ("emission", C.c_float * 3),
        ("roughness", C.c_float), ("metallic", C.c_float), ("transmission", C.c_float), ("ior", C.c_float),
    ]


class RtpbrCamera(C.Structure):
    _fields_ = [
        ("lookfrom", C.c_float * 3), ("lookat", C.c_float * 3), ("vup", C.c_float * 3),
        ("vfov", C.c_float), ("aspect", C.c_float), ("aperture", C.c_float), ("focus", C.c_float),
    ]


class RtpbrConfig(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32),
        ("family", C.c_int32), ("max_bounces", C.c_int32), ("max_steps", C.c_int32), ("marcher", C.c_int32),
        ("t_start", C.c_float), ("hit_eps", C.c_float), ("t_far", C.c_float),
        ("relax_w0", C.c_float), ("relax_guard", C.c_int32), ("relax_reset", C.c_int32), ("relax_w_reset", C.c_float),
        ("normal_h", C.c_float), ("box_round", C.c_float), ("light_quality", C.c_float),
        ("bsdf", C.c_int32), ("f0_variant", C.c_int32),
        ("visibility_min", C.c_float), ("visibility_max", C.c_float),
        ("sky", C.c_int32), ("sky_scale", C.c_float),
        ("seed", C.c_uint32),
        ("min_dis", C.c_float), ("pixel_radius", C.c_float), ("quality_per_sample", C.c_float),
        ("black_background", C.c_int32),
        ("nearest_seed", C.c_int32), ("normal_mode", C.c_int32), ("samples_per_pixel", C.c_int32),
        ("adaptive_sampling", C.c_int32), ("noise_threshold", C.c_float),
        ("inner_spp", C.c_int32), ("primary_miss", C.c_int32), ("bunny_bob", C.c_int32),
        ("kernel", C.c_int32), ("count_work", C.c_int32),
    ]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile librtpbr.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    srcs.append(os.path.join(_HERE, "..", "include", "rtpbr.h"))
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        cmd = ["make", "-C", CSRC] + (["-B"] if force else [])
        subprocess.run(cmd, check=True, stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """Load librtpbr.so; fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C raytracingpbr_b200/csrc`.  There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    sig = {
        "rtpbr_create": [C.POINTER(RtpbrConfig), C.c_int, C.POINTER(vp)],
        "rtpbr_destroy": [vp],
        "rtpbr_set_scene": [vp, C.POINTER(RtpbrObject), C.c_int],
        "rtpbr_set_camera": [vp, C.POINTER(RtpbrCamera)],
        "rtpbr_set_envmap": [vp, C.POINTER(C.c_float), C.c_int, C.c_int],
        "rtpbr_set_frame": [vp, C.c_int],
        "rtpbr_set_sample_base": [vp, C.c_uint32],
        "rtpbr_set_shard": [vp, C.c_int, C.c_int, C.c_int],
        "rtpbr_refresh": [vp],
        "rtpbr_pathtrace": [vp, C.c_int],
        "rtpbr_post_process": [vp, C.c_int, C.c_float, C.c_double],
        "rtpbr_download": [vp, C.c_int, vp, C.c_size_t],
        "rtpbr_upload": [vp, C.c_int, vp, C.c_size_t],
        "rtpbr_sync": [vp],
        "rtpbr_flush_l2": [vp],
        "rtpbr_alloc_host": [C.c_size_t, C.POINTER(vp)],
        "rtpbr_free_host": [vp],
        "rtpbr_timer_start": [vp],
        "rtpbr_timer_stop": [vp, C.POINTER(C.c_float)],
        "rtpbr_kernel_time": [vp, C.POINTER(C.c_float), C.POINTER(C.c_int)],
        "rtpbr_get_counters": [vp, C.POINTER(C.c_uint64)],
        "rtpbr_device_info": [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)],
        "rtpbr_nccl_unique_id": [vp],
        "rtpbr_nccl_init": [vp, vp, C.c_int, C.c_int],
        "rtpbr_reduce_tiles": [vp, C.c_int],
        "rtpbr_device_ptr": [vp, C.c_int, C.POINTER(C.c_uint64)],
        "rtpbr_set_jit": [vp, C.c_int],
        "rtpbr_jit_status": [vp, C.c_char_p, C.c_size_t],
        "rtpbr_jit_compile_check": [C.POINTER(RtpbrConfig), C.POINTER(RtpbrObject), C.c_int, C.c_char_p, C.c_size_t],
        "rtpbr_version": [], "rtpbr_sizeof_config": [], "rtpbr_sizeof_object": [], "rtpbr_sizeof_camera": [],
        "rtpbr_device_count": [],
        "rtpbr_denoise": [vp, C.c_float],
        "rtpbr_multi_create": [C.POINTER(RtpbrConfig), C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(vp)],
        "rtpbr_multi_destroy": [vp],
        "rtpbr_multi_count": [vp],
        "rtpbr_multi_set_scene": [vp, C.POINTER(RtpbrObject), C.c_int],
        "rtpbr_multi_set_camera": [vp, C.POINTER(RtpbrCamera)],
        "rtpbr_multi_set_envmap": [vp, C.POINTER(C.c_float), C.c_int, C.c_int],
        "rtpbr_multi_set_frame": [vp, C.c_int],
        "rtpbr_multi_set_sample_base": [vp, C.c_uint32],
        "rtpbr_multi_refresh": [vp],
        "rtpbr_multi_pathtrace": [vp, C.c_int],
        "rtpbr_multi_reduce": [vp, C.c_int],
        "rtpbr_multi_post_process": [vp, C.c_int, C.c_float, C.c_double],
        "rtpbr_multi_download": [vp, C.c_int, vp, C.c_size_t],
        "rtpbr_multi_sync": [vp],
    }
    for name, argtypes in sig.items():
        fn = getattr(L, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    L.rtpbr_jit_generate.argtypes = [C.POINTER(RtpbrConfig), C.POINTER(RtpbrObject), C.c_int, C.c_char_p, C.c_size_t]
    L.rtpbr_jit_generate.restype = C.c_longlong
    L.rtpbr_multi_context.argtypes = [vp, C.c_int]
    L.rtpbr_multi_context.restype = vp
    L.rtpbr_last_error.argtypes = []
    L.rtpbr_last_error.restype = C.c_char_p
    if L.rtpbr_sizeof_config() != C.sizeof(RtpbrConfig) or L.rtpbr_sizeof_object() != C.sizeof(RtpbrObject) \
            or L.rtpbr_sizeof_camera() != C.sizeof(RtpbrCamera):
        raise ImportError("librtpbr.so struct layout differs from the ctypes mirror: rebuild the library")
    _lib = L
    return L


def device_count() -> int:
    return int(lib().rtpbr_device_count())


def check(rc: int) -> None:
    if rc != 0:
        raise RtpbrError(rc, (lib().rtpbr_last_error() or b"").decode("utf-8", "replace"))


class PinnedArray:
    """Page-locked host array (cudaHostAlloc) exposed as a numpy view; free() or use as a context manager."""

    def __init__(self, shape, dtype=np.float32):
        self._L = lib()
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        check(self._L.rtpbr_alloc_host(self.nbytes, C.byref(p)))
        self._p = p
        buf = (C.c_char * self.nbytes).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def free(self):
        if getattr(self, "_p", None):
            self.array = None
            self._L.rtpbr_free_host(self._p)
            self._p = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.free()


def jit_source(cfg: RtpbrConfig, objects) -> str:
    """The scene-specialised translation unit rtpbr_pathtrace would compile (no GPU needed)."""
    arr = (RtpbrObject * len(objects))(*objects)
    n = lib().rtpbr_jit_generate(C.byref(cfg), arr, len(objects), None, 0)
    if n < 0:
        check(int(n))
    buf = C.create_string_buffer(int(n) + 1)
    lib().rtpbr_jit_generate(C.byref(cfg), arr, len(objects), buf, int(n) + 1)
    return buf.value.decode()


def jit_compile_check(cfg: RtpbrConfig, objects) -> str:
    """Run NVRTC on the specialised source (no GPU needed); returns the compiler log, raises on failure."""
    _point_at_nvrtc()
    arr = (RtpbrObject * len(objects))(*objects)
    log = C.create_string_buffer(1 << 16)
    check(lib().rtpbr_jit_compile_check(C.byref(cfg), arr, len(objects), log, len(log)))
    return log.value.decode()


def _point_at_nvrtc() -> None:
    """Help librtpbr find the pip-installed NVRTC when the CUDA toolkit's is not on the loader path."""
    if os.environ.get("RTPBR_NVRTC_LIB"):
        return
    import importlib.util
    spec = importlib.util.find_spec("nvidia")
    if spec is None or not spec.submodule_search_locations:
        return
    for base in spec.submodule_search_locations:
        p = os.path.join(base, "cuda_nvrtc", "lib", "libnvrtc.so.12")
        if os.path.exists(p) and not os.path.exists("/usr/local/cuda/lib64/libnvrtc.so.12"):
            os.environ["RTPBR_NVRTC_LIB"] = p
            return


def _find_nccl() -> str | None:
    """Locate the pip-installed NCCL (nvidia-nccl-cu12) without importing torch."""
    import importlib.util
    spec = importlib.util.find_spec("nvidia")
    if spec is None or not spec.submodule_search_locations:
        return None
    for base in spec.submodule_search_locations:
        p = os.path.join(base, "nccl", "lib", "libnccl.so.2")
        if os.path.exists(p):
            return p
    return None


class Context:
    """Owns one RtpbrContext (one GPU, one stream, all device buffers)."""

    def __init__(self, cfg: RtpbrConfig, device: int = 0, _borrowed=None):
        self._L = lib()
        self.cfg = cfg
        self.width, self.height = cfg.width, cfg.height
        self._owned = _borrowed is None
        if _borrowed is not None:          # a context lent by a MultiContext (rtpbr_multi_context)
            self._h = _borrowed
            return
        h = C.c_void_p()
        _point_at_nvrtc()
        check(self._L.rtpbr_create(C.byref(cfg), device, C.byref(h)))
        self._h = h

    def close(self) -> None:
        if getattr(self, "_h", None):
            if self._owned:
                self._L.rtpbr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- scene / camera ---------------------------------------------------------------
    def set_scene(self, objects) -> None:
        arr = (RtpbrObject * len(objects))(*objects)
        check(self._L.rtpbr_set_scene(self._h, arr, len(objects)))

    def set_camera(self, cam: RtpbrCamera) -> None:
        check(self._L.rtpbr_set_camera(self._h, C.byref(cam)))

    def set_envmap(self, rgb: np.ndarray) -> None:
        rgb = np.ascontiguousarray(rgb, dtype=np.float32)
        assert rgb.ndim == 3 and rgb.shape[2] == 3
        check(self._L.rtpbr_set_envmap(self._h, rgb.ctypes.data_as(C.POINTER(C.c_float)), rgb.shape[0], rgb.shape[1]))

    def set_frame(self, frame: int) -> None:
        check(self._L.rtpbr_set_frame(self._h, int(frame)))

    def set_sample_base(self, base: int) -> None:
        check(self._L.rtpbr_set_sample_base(self._h, int(base)))

    def set_shard(self, rank: int, nranks: int, band: int = 32) -> None:
        check(self._L.rtpbr_set_shard(self._h, rank, nranks, band))

    # -- kernels ----------------------------------------------------------------------
    def refresh(self) -> None:
        check(self._L.rtpbr_refresh(self._h))

    def pathtrace(self, spp: int = 1) -> None:
        check(self._L.rtpbr_pathtrace(self._h, int(spp)))

    def post_process(self, mode: int, exposure: float = 1.0, gamma: float = 2.2) -> None:
        check(self._L.rtpbr_post_process(self._h, mode, exposure, gamma))

    def denoise(self, threshold: float) -> None:
        check(self._L.rtpbr_denoise(self._h, float(threshold)))

    def sync(self) -> None:
        check(self._L.rtpbr_sync(self._h))

    def set_jit(self, enable: bool) -> None:
        check(self._L.rtpbr_set_jit(self._h, int(enable)))

    def jit_status(self):
        """(active, description) of the scene-specialised kernel."""
        buf = C.create_string_buffer(4096)
        rc = self._L.rtpbr_jit_status(self._h, buf, len(buf))
        if rc < 0:
            check(rc)
        return bool(rc), buf.value.decode("utf-8", "replace")

    def flush_l2(self) -> None:
        check(self._L.rtpbr_flush_l2(self._h))

    # -- data -------------------------------------------------------------------------
    def download(self, which: int = BUF_IMAGE_BUFFER, out: np.ndarray | None = None) -> np.ndarray:
        ch = {BUF_IMAGE_BUFFER: 4, BUF_IMAGE_PIXELS: 3, BUF_RAY_BUFFER: 10, BUF_DIFF_BUFFER: 2, BUF_DIFF_PIXELS: 1, BUF_DENOISE_PIXELS: 3}[which]
        if out is None:
            out = np.empty((self.width, self.height, ch), dtype=np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.shape == (self.width, self.height, ch)
        check(self._L.rtpbr_download(self._h, which, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def upload(self, which: int, arr: np.ndarray) -> None:
        arr = np.ascontiguousarray(arr, dtype=np.float32)
        check(self._L.rtpbr_upload(self._h, which, arr.ctypes.data_as(C.c_void_p), arr.nbytes))

    def device_ptr(self, which: int = BUF_IMAGE_BUFFER) -> int:
        p = C.c_uint64()
        check(self._L.rtpbr_device_ptr(self._h, which, C.byref(p)))
        return p.value

    # -- measurement ------------------------------------------------------------------
    def timer_start(self) -> None:
        check(self._L.rtpbr_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        check(self._L.rtpbr_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def kernel_time(self):
        ms, n = C.c_float(), C.c_int()
        check(self._L.rtpbr_kernel_time(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def counters(self) -> dict:
        out = (C.c_uint64 * len(CNT_NAMES))()
        check(self._L.rtpbr_get_counters(self._h, out))
        return dict(zip(CNT_NAMES, [int(v) for v in out]))

    def device_info(self) -> dict:
        a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        check(self._L.rtpbr_device_info(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return {"sm_count": a.value, "cc": (b.value, c.value), "blocks_per_sm": d.value}

    # -- multi-GPU --------------------------------------------------------------------
    @staticmethod
    def nccl_unique_id() -> bytes:
        p = _find_nccl()
        if p and not os.environ.get("RTPBR_NCCL_LIB"):
            os.environ["RTPBR_NCCL_LIB"] = p
        buf = C.create_string_buffer(128)
        check(lib().rtpbr_nccl_unique_id(buf))
        return buf.raw

    def nccl_init(self, unique_id: bytes, rank: int, nranks: int) -> None:
        p = _find_nccl()
        if p and not os.environ.get("RTPBR_NCCL_LIB"):
            os.environ["RTPBR_NCCL_LIB"] = p
        assert len(unique_id) == 128
        buf = C.create_string_buffer(unique_id, 128)
        check(self._L.rtpbr_nccl_init(self._h, buf, rank, nranks))

    def reduce_tiles(self, root: int = 0) -> None:
        check(self._L.rtpbr_reduce_tiles(self._h, root))


class MultiContext:
    """Owns one RtpbrMulti: n GPUs driven by this one process (rtpbr_multi_*, include/rtpbr.h).  No torch, no MPI: the
    NCCL communicators are created by the library itself."""

    def __init__(self, cfg: RtpbrConfig, devices, band: int = 4):
        self._L = lib()
        self.cfg = cfg
        self.width, self.height = cfg.width, cfg.height
        devices = list(devices)
        _point_at_nvrtc()
        p = _find_nccl()
        if p and not os.environ.get("RTPBR_NCCL_LIB"):
            os.environ["RTPBR_NCCL_LIB"] = p
        h = C.c_void_p()
        arr = (C.c_int * len(devices))(*devices)
        check(self._L.rtpbr_multi_create(C.byref(cfg), arr, len(devices), band, C.byref(h)))
        self._h = h
        self.devices, self.band = devices, band
        self.ranks = [Context(cfg, _borrowed=C.c_void_p(self._L.rtpbr_multi_context(h, r))) for r in range(len(devices))]

    def close(self) -> None:
        if getattr(self, "_h", None):
            for c in self.ranks:
                c.close()
            self._L.rtpbr_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_scene(self, objects) -> None:
        arr = (RtpbrObject * len(objects))(*objects)
        check(self._L.rtpbr_multi_set_scene(self._h, arr, len(objects)))

    def set_camera(self, cam: RtpbrCamera) -> None:
        check(self._L.rtpbr_multi_set_camera(self._h, C.byref(cam)))

    def set_envmap(self, rgb: np.ndarray) -> None:
        rgb = np.ascontiguousarray(rgb, dtype=np.float32)
        assert rgb.ndim == 3 and rgb.shape[2] == 3
        check(self._L.rtpbr_multi_set_envmap(self._h, rgb.ctypes.data_as(C.POINTER(C.c_float)), rgb.shape[0], rgb.shape[1]))

    def set_frame(self, frame: int) -> None:
        check(self._L.rtpbr_multi_set_frame(self._h, int(frame)))

    def set_sample_base(self, base: int) -> None:
        check(self._L.rtpbr_multi_set_sample_base(self._h, int(base)))

    def refresh(self) -> None:
        check(self._L.rtpbr_multi_refresh(self._h))

    def pathtrace(self, spp: int = 1) -> None:
        check(self._L.rtpbr_multi_pathtrace(self._h, int(spp)))

    def reduce(self, root: int = 0) -> None:
        check(self._L.rtpbr_multi_reduce(self._h, root))

    def post_process(self, mode: int, exposure: float = 1.0, gamma: float = 2.2) -> None:
        check(self._L.rtpbr_multi_post_process(self._h, mode, exposure, gamma))

    def sync(self) -> None:
        check(self._L.rtpbr_multi_sync(self._h))

    def download(self, which: int = BUF_IMAGE_BUFFER, out: np.ndarray | None = None) -> np.ndarray:
        ch = {BUF_IMAGE_BUFFER: 4, BUF_IMAGE_PIXELS: 3, BUF_RAY_BUFFER: 10, BUF_DIFF_BUFFER: 2, BUF_DIFF_PIXELS: 1, BUF_DENOISE_PIXELS: 3}[which]
        if out is None:
            out = np.empty((self.width, self.height, ch), dtype=np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.shape == (self.width, self.height, ch)
        check(self._L.rtpbr_multi_download(self._h, which, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def jit_status(self):
        return self.ranks[0].jit_status()
