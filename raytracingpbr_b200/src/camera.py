"""src/camera.py -- camera state.  `get_ray` runs inside the CUDA kernel (rt_integrator.cuh
camera_ray); the easing of SmoothCamera (src/camera.py:82-112)
is restated on the host."""
import numpy as np

from . import config
from ._runtime import ScalarField


class SmoothCamera:
    """src/camera.py:39-112.  The easing kernel `_update` (:82-112) is a few scalar operations per frame, so it
    runs on the host in fp32; `_rotate` needs the GGUI camera helpers and is out of scope."""

    def __init__(self):
        self.position = ScalarField(np.array([0.0, -0.2, 4.0], dtype=np.float32))   # src/main.py:17
        self.lookat = ScalarField(np.array([0.0, -0.2, 3.0], dtype=np.float32))
        self.up = ScalarField(np.array([0.0, 1.0, 0.0], dtype=np.float32))
        self.position_velocity = ScalarField(np.float32(10), camera=False)           # :53-55
        self.lookat_velocity = ScalarField(np.float32(10), camera=False)
        self.up_velocity = ScalarField(np.float32(10), camera=False)
        self.moving = ScalarField(0, camera=False)

    def init(self, camera):
        """camera: anything with curr_position / curr_lookat / curr_up (ti.ui.Camera surface); :58-61."""
        self.position[None] = np.asarray(camera.curr_position, dtype=np.float32)
        self.lookat[None] = np.asarray(camera.curr_lookat, dtype=np.float32)
        self.up[None] = np.asarray(camera.curr_up, dtype=np.float32)

    def update(self, dt, camera, direction=None):
        """:63-66 without _rotate (GUI)."""
        self._update(dt, camera.curr_position, camera.curr_lookat, camera.curr_up)

    def _update(self, dt, curr_position, curr_lookat, curr_up):
        """:82-112: exponential easing towards the GUI camera; sets `moving`, bumps u_frame."""
        from .fileds import u_frame
        f = np.float32
        dt = f(dt)
        diffs = []
        for fld, vel, cur in ((self.position, self.position_velocity, curr_position),
                              (self.lookat, self.lookat_velocity, curr_lookat), (self.up, self.up_velocity, curr_up)):
            val = fld[None]
            diff = np.asarray(cur, dtype=np.float32) - val
            fld[None] = (val + diff * np.clip(vel[None] * dt, f(0), f(1))).astype(np.float32)
            diffs.append(float(np.abs(diff).max()))
        self.moving[None] = int(max(diffs) > 1e-3)
        u_frame[None] = u_frame[None] + 1


smooth = SmoothCamera()
camera_gamma = 2.2


def _f32(x):
    return np.float32(x)


# src/camera.py:125-129
aspect_ratio = ScalarField(float((_f32(1.0) / _f32(config.image_resolution[1])) / (_f32(1.0) / _f32(config.image_resolution[0]))))
camera_exposure = ScalarField(1.0, camera=False)
camera_vfov = ScalarField(35.0)
camera_aperture = ScalarField(0.01)
camera_focus = ScalarField(4.0)
