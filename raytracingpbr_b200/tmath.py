"""Minimal host-side stand-ins for the `taichi.math` constructors the reference uses to
*describe* scenes (vec2/vec3/vec4, broadcasting, fp32 arithmetic).  Device math lives in
csrc/rt_math.cuh; nothing here is on the hot path."""
from __future__ import annotations

import math

import numpy as np

pi = math.pi


def _vec(n, args):
    if len(args) == 0:
        return np.zeros(n, dtype=np.float32)
    if len(args) == 1 and np.ndim(args[0]) == 0:
        return np.full(n, args[0], dtype=np.float32)          # vec3(x) broadcast
    flat = np.concatenate([np.atleast_1d(np.asarray(a, dtype=np.float32)) for a in args])
    if flat.shape != (n,):
        raise ValueError(f"vec{n} needs {n} components, got {flat.shape[0]}")
    return flat.astype(np.float32)


def vec2(*args):
    return _vec(2, args)


def vec3(*args):
    return _vec(3, args)


def vec4(*args):
    return _vec(4, args)


def radians(deg):
    return np.float32(deg) * np.float32(math.pi / 180.0)
