"""Lazy singleton behind the `src`-shaped modules: owns the PathTracer of the family-C integrator."""
from __future__ import annotations

import numpy as np

from .. import _native as N
from .. import scenes
from ..dataclass import Camera
from ..engine import PathTracer
from ..tmath import vec3
from . import config

_pt: PathTracer | None = None
_dirty_camera = True
_dirty_scene = True
_env: np.ndarray | None = None
_dirty_env = False


class ScalarField:
    """`ti.field(dtype, shape=())`: `f[None]` reads / writes (src/camera.py:117-129, src/fileds.py:15)."""
    def __init__(self, value, camera=True):
        self._v, self._camera = value, camera

    def __getitem__(self, key):
        return self._v

    def __setitem__(self, key, value):
        global _dirty_camera
        self._v = value
        if self._camera:
            _dirty_camera = True


def mark_scene_dirty():
    global _dirty_scene
    _dirty_scene = True


def set_env(table):
    global _env, _dirty_env
    _env, _dirty_env = np.ascontiguousarray(table, dtype=np.float32), True


def tracer() -> PathTracer:
    """Create (first call) or refresh the GPU context from the module-level state."""
    global _pt, _dirty_camera, _dirty_scene, _dirty_env
    from . import camera as cam
    from . import scene
    if _pt is None:
        w, h = config.image_resolution
        cfg, _, _, tm = scenes.src_scene(w, h, max_bounces=config.MAX_RAYTRACE, seed=config.SEED)
        cfg.max_steps = config.MAX_RAYMARCH
        cfg.quality_per_sample = config.QUALITY_PER_SAMPLE
        cfg.black_background = int(config.BLACK_BACKGROUND)
        cfg.visibility_min, cfg.visibility_max = config.VISIBILITY
        cfg.samples_per_pixel = config.SAMPLES_PER_PIXEL
        cfg.adaptive_sampling, cfg.noise_threshold = int(config.ADAPTIVE_SAMPLING), config.NOISE_THRESHOLD
        _pt = PathTracer(cfg, scene.OBJECTS, _camera(cam), tm, device=config.DEVICE)
        _dirty_camera = _dirty_scene = False
        _dirty_env = _env is not None
    if _dirty_scene:
        _pt.set_scene(scene.OBJECTS)
        _dirty_scene = False
    if _dirty_camera:
        _pt.set_camera(_camera(cam))
        _dirty_camera = False
    if _dirty_env:
        _pt.set_envmap(_env)
        _dirty_env = False
    _pt.tonemap["exposure"] = float(cam.camera_exposure[None])
    return _pt


def _camera(cam) -> Camera:
    return Camera(vec3(*cam.smooth.position[None]), vec3(*cam.smooth.lookat[None]), vec3(*cam.smooth.up[None]),
                  float(cam.camera_vfov[None]), float(cam.aspect_ratio[None]), float(cam.camera_aperture[None]),
                  float(cam.camera_focus[None]))


def close():
    global _pt
    if _pt is not None:
        _pt.close()
        _pt = None
