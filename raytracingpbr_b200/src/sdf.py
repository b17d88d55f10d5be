"""src/sdf.py:12-18 -- the SHAPE enum; the primitives themselves live in csrc/rt_integrator.cuh."""
from enum import IntEnum


class SHAPE(IntEnum):
    NONE = 0
    SPHERE = 1
    BOX = 2
    CYLINDER = 3
    CONE = 4
    PLANE = 5
