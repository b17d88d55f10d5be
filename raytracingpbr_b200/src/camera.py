"""src/camera.py -- camera state.  `get_ray` runs inside the CUDA kernel (rt_integrator.cuh
camera_ray); the GUI easing of SmoothCamera (src/camera.py:82-112) is out of scope: update()
jumps straight to the target."""
import numpy as np

from . import config
from ._runtime import ScalarField


class SmoothCamera:
    def __init__(self):
        self.position = ScalarField(np.array([0.0, -0.2, 4.0], dtype=np.float32))   # src/main.py:17
        self.lookat = ScalarField(np.array([0.0, -0.2, 3.0], dtype=np.float32))
        self.up = ScalarField(np.array([0.0, 1.0, 0.0], dtype=np.float32))
        self.moving = ScalarField(0, camera=False)

    def init(self, camera):
        """camera: anything with curr_position / curr_lookat / curr_up (ti.ui.Camera surface)."""
        self.position[None] = np.asarray(camera.curr_position, dtype=np.float32)
        self.lookat[None] = np.asarray(camera.curr_lookat, dtype=np.float32)
        self.up[None] = np.asarray(camera.curr_up, dtype=np.float32)

    def update(self, dt, camera, direction=None):
        before = (self.position[None].copy(), self.lookat[None].copy(), self.up[None].copy())
        self.init(camera)
        after = (self.position[None], self.lookat[None], self.up[None])
        self.moving[None] = int(any(np.abs(a - b).max() > 1e-3 for a, b in zip(after, before)))


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
