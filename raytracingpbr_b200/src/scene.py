"""src/scene.py -- the object table.  `nearest` / `raycast` / `calc_normal` run inside the CUDA
kernel; `build_scene()` (src/scene.py:112-113) uploads OBJECTS, deriving Transform.matrix."""
from .. import scenes
from . import _runtime

OBJECTS = scenes._demo_objects("src")     # src/scene.py:11-33, sorted by type
objects = OBJECTS                          # the reference's field mirrors the list


def build_scene():
    _runtime.mark_scene_dirty()
