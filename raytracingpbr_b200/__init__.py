"""raytracingpbr_b200 -- B200-native SDF path tracer with the Python surface of
HK-SHAO/RayTracingPBR.  The hot path is a hand-written sm_100a CUDA kernel reached through the
C-ABI in include/rtpbr.h (ctypes; no Taichi, no Triton, no PyTorch).  No CPU fallback."""
from . import _native
from ._native import RtpbrError, build
from .dataclass import Camera, Material, Ray, SDFObject, Transform
from .engine import MultiPathTracer, PathTracer, imwrite

nccl_unique_id = _native.Context.nccl_unique_id
from .tmath import vec2, vec3, vec4

__all__ = ["_native", "RtpbrError", "build", "Camera", "Material", "Ray", "SDFObject", "Transform", "PathTracer", "MultiPathTracer", "nccl_unique_id",
           "imwrite", "vec2", "vec3", "vec4"]
