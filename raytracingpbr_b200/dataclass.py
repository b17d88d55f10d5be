"""The reference's dataclass surface (src/dataclass.py:5-46), as plain Python dataclasses.

Construction is positional with trailing defaults exactly like the reference's
``@ti.dataclass`` structs, e.g. ``Transform(vec3(0, 0, -1), vec3(0), vec3(1, 1, 0.2))`` and
``Material(vec3(1, 1, 1)*0.6, vec3(1), 1.0, 1.0, 0, 1.100)`` (src/scene.py:13-14).
``Transform.matrix`` is derived on upload (rtpbr_set_scene), as the reference does with its
``update_all_transform`` kernel (src/scene.py:99-109).
"""
from __future__ import annotations

import dataclasses

import numpy as np

from . import _native as N
from .tmath import vec3


def _v3(x):
    return np.asarray(x, dtype=np.float32).reshape(3).copy()


def _field(default):
    return dataclasses.field(default_factory=lambda: vec3(default))


@dataclasses.dataclass
class Ray:                       # src/dataclass.py:5-10
    origin: np.ndarray = _field(0)
    direction: np.ndarray = _field(0)
    color: np.ndarray = _field(0)
    depth: int = 0


@dataclasses.dataclass
class Material:                  # src/dataclass.py:13-20
    albedo: np.ndarray = _field(0)
    emission: np.ndarray = _field(0)
    roughness: float = 0.0
    metallic: float = 0.0
    transmission: float = 0.0
    ior: float = 0.0


@dataclasses.dataclass
class Transform:                 # src/dataclass.py:23-28
    position: np.ndarray = _field(0)
    rotation: np.ndarray = _field(0)
    scale: np.ndarray = _field(0)
    matrix: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros((3, 3), dtype=np.float32))


@dataclasses.dataclass
class SDFObject:                 # src/dataclass.py:31-35 (+ `distance` of the examples, cornell_box.py:62-67)
    type: int = 0
    transform: Transform = dataclasses.field(default_factory=Transform)
    material: Material = dataclasses.field(default_factory=Material)
    distance: float = 0.0

    def to_native(self) -> N.RtpbrObject:
        o = N.RtpbrObject()
        o.type = int(self.type)
        o.position[:] = _v3(self.transform.position).tolist()
        o.rotation[:] = _v3(self.transform.rotation).tolist()
        o.scale[:] = _v3(self.transform.scale).tolist()
        o.albedo[:] = _v3(self.material.albedo).tolist()
        o.emission[:] = _v3(self.material.emission).tolist()
        o.roughness = float(self.material.roughness)
        o.metallic = float(self.material.metallic)
        o.transmission = float(self.material.transmission)
        o.ior = float(self.material.ior)
        return o


@dataclasses.dataclass
class Camera:                    # src/dataclass.py:38-46
    lookfrom: np.ndarray = _field(0)
    lookat: np.ndarray = _field(0)
    vup: np.ndarray = _field(0)
    vfov: float = 0.0
    aspect: float = 0.0
    aperture: float = 0.0
    focus: float = 0.0

    def to_native(self) -> N.RtpbrCamera:
        c = N.RtpbrCamera()
        c.lookfrom[:] = _v3(self.lookfrom).tolist()
        c.lookat[:] = _v3(self.lookat).tolist()
        c.vup[:] = _v3(self.vup).tolist()
        c.vfov, c.aspect, c.aperture, c.focus = float(self.vfov), float(self.aspect), float(self.aperture), float(self.focus)
        return c
