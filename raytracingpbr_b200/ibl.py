"""Environment-map input side of the hot path ("next" row of SURVEY.md 8(f)): what the reference
does in `Image.__init__` / `Image.process` (src/ibl.py:12-23, tokyo_ibl.py:40-51,
bunny_sdf_glass.py:41-45, 277-281) before the kernel's nearest-texel lookup.

The reference loads its `.hdr` files with `ti.tools.imread`, i.e. through stb_image's 8-bit
path [TAICHI-INTERNAL, SURVEY.md 8(c)]: Radiance RGBE -> float -> LDR `clamp(x^(1/2.2)*255+0.5)`
-> uint8 (W, H, 3) with y pointing up.  It then re-linearises: `(u8 / 255 * exposure) ** gamma`.
This module restates that pipeline in numpy (host side, once per scene); the resulting table is
what `PathTracer.set_envmap` uploads.  pow() follows the fp32 contract: binary64 libm, rounded once.
"""
from __future__ import annotations

import numpy as np


def read_rgbe(path: str) -> np.ndarray:
    """Decode a Radiance `.hdr` (RGBE, flat or new-style RLE scanlines) to float32 (H, W, 3),
    first row = top of the image.  Mantissa scaling as in stb_image: m * 2^(e - 136), e = 0 -> 0."""
    with open(path, "rb") as f:
        data = f.read()
    pos = 0
    if not (data.startswith(b"#?RADIANCE") or data.startswith(b"#?RGBE")):
        raise ValueError(f"{path}: not a Radiance HDR file")
    fmt_ok = False
    while True:
        end = data.index(b"\n", pos)
        line = data[pos:end]
        pos = end + 1
        if line == b"":
            break
        if line.startswith(b"FORMAT=") and b"32-bit_rle_rgbe" in line:
            fmt_ok = True
    if not fmt_ok:
        raise ValueError(f"{path}: unsupported FORMAT")
    end = data.index(b"\n", pos)
    dims = data[pos:end].split()
    pos = end + 1
    if len(dims) != 4 or dims[0] != b"-Y" or dims[2] != b"+X":
        raise ValueError(f"{path}: unsupported orientation {dims}")
    h, w = int(dims[1]), int(dims[3])
    buf = np.frombuffer(data, dtype=np.uint8, offset=pos)
    rgbe = np.empty((h, w, 4), dtype=np.uint8)
    p = 0
    if w < 8 or w >= 32768 or not (buf[0] == 2 and buf[1] == 2 and not (buf[2] & 0x80)):
        rgbe[:] = buf[:h * w * 4].reshape(h, w, 4)            # flat
    else:
        for y in range(h):
            if buf[p] != 2 or buf[p + 1] != 2 or ((int(buf[p + 2]) << 8) | int(buf[p + 3])) != w:
                raise ValueError(f"{path}: bad scanline header at row {y}")
            p += 4
            for c in range(4):
                x = 0
                row = rgbe[y, :, c]
                while x < w:
                    n = int(buf[p]); p += 1
                    if n > 128:                               # run
                        n -= 128
                        row[x:x + n] = buf[p]; p += 1
                    else:                                     # literal
                        row[x:x + n] = buf[p:p + n]; p += n
                    x += n
    e = rgbe[..., 3].astype(np.int32)
    scale = np.where(e == 0, np.float32(0), np.ldexp(np.float32(1.0), e - 136).astype(np.float32))
    return rgbe[..., :3].astype(np.float32) * scale[..., None]


def hdr_to_ldr_stb(rgb: np.ndarray) -> np.ndarray:
    """stb_image's HDR -> LDR conversion (stbi__hdr_to_ldr, gamma 2.2, scale 1):
    z = pow(x, 1/2.2) * 255 + 0.5, clamped to [0, 255], truncated."""
    z = np.power(np.maximum(rgb.astype(np.float32), 0), np.float32(1.0 / 2.2)) * np.float32(255) + np.float32(0.5)
    return np.clip(z, 0, 255).astype(np.uint8)


def imread(path: str) -> np.ndarray:
    """What `ti.tools.imread(path)` returns for the reference's `.hdr` assets: uint8 (W, H, 3),
    second axis pointing up (src/ibl.py:15)."""
    ldr = hdr_to_ldr_stb(read_rgbe(path))          # (H, W, 3), top row first
    return np.ascontiguousarray(ldr.swapaxes(0, 1)[:, ::-1, :])


def process(u8: np.ndarray, exposure: float, gamma: float) -> np.ndarray:
    """`img / 255` (src/ibl.py:17) then `Image.process(exposure, gamma)` (src/ibl.py:19-23 with
    adjust(), src/postprocessor.py:17-21; tokyo_ibl.py:46-51): (c * exposure) ** gamma.
    For bunny_sdf_glass.py pass exposure 1.8, gamma 2.2: it applies the same map per lookup (:279-280)."""
    x = (np.asarray(u8).astype(np.float32) / np.float32(255)) * np.float32(exposure)
    return np.ascontiguousarray(np.power(x.astype(np.float64), float(np.float32(gamma))).astype(np.float32))


def load_envmap(path: str, exposure: float, gamma: float = 2.2) -> np.ndarray:
    """`.hdr` file -> processed (W, H, 3) f32 table for PathTracer.set_envmap."""
    return process(imread(path), exposure, gamma)


def write_rgbe(path: str, rgb: np.ndarray) -> None:
    """Minimal flat (non-RLE) Radiance writer; used by tests and tools to make small fixtures."""
    rgb = np.maximum(np.asarray(rgb, dtype=np.float32), 0)
    h, w, _ = rgb.shape
    m = rgb.max(axis=2)
    mant, ex = np.frexp(m)                                    # m = mant * 2^ex, mant in [0.5, 1)
    scale = np.where(m < 1e-32, 0.0, 256.0 / np.ldexp(1.0, ex))
    out = np.zeros((h, w, 4), dtype=np.uint8)
    out[..., :3] = np.clip(rgb * scale[..., None], 0, 255).astype(np.uint8)
    out[..., 3] = np.where(m < 1e-32, 0, ex + 128).astype(np.uint8)
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n")
        f.write(f"-Y {h} +X {w}\n".encode())
        f.write(out.tobytes())
