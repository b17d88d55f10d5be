"""src/dataclass.py:5-46"""
from ..dataclass import Camera, Material, Ray, SDFObject, Transform  # noqa: F401
