"""taichi.ui stand-in: a window that never runs, so example scripts fall through their GUI loop."""
from ..math import vec3

LEFT, RIGHT, UP, DOWN, RELEASE, LMB = "Left", "Right", "Up", "Down", "Release", "LMB"


class _Canvas:
    def set_image(self, *a, **kw):
        return None


class Window:
    running = False

    def __init__(self, *a, **kw):
        pass

    def get_canvas(self):
        return _Canvas()

    def is_pressed(self, *a):
        return False

    def get_events(self, *a):
        return []

    def show(self):
        return None


class Camera:
    def __init__(self):
        self.curr_position = vec3(0, 0, 0)
        self.curr_lookat = vec3(0, 0, 1)     # [TAICHI-INTERNAL] default look-at; generators pass explicit cameras
        self.curr_up = vec3(0, 1, 0)

    def position(self, x, y, z):
        d = self.curr_lookat - self.curr_position
        self.curr_position = vec3(x, y, z)
        self.curr_lookat = self.curr_position + d

    def lookat(self, x, y, z):
        self.curr_lookat = vec3(x, y, z)

    def up(self, x, y, z):
        self.curr_up = vec3(x, y, z)

    def track_user_inputs(self, *a, **kw):
        return None
