"""PathTracer: the host-side object a user drives (one GPU).  Mirrors the call sequence of the
reference's frame loops -- refresh() / pathtrace() x N / post_process() (src/renderer.py:25-32,
bunny_sdf_glass.py:437-451) -- over the C-ABI in include/rtpbr.h."""
from __future__ import annotations

import numpy as np

from . import _native as N
from .dataclass import Camera, SDFObject


class Field:
    """Read-only view with the `.to_numpy()` / `.shape` surface of a Taichi field
    (src/fileds.py:7-13): dense (W, H, C) float32, j fastest, origin bottom-left."""

    def __init__(self, ctx: N.Context, which: int, channels: int):
        self._ctx, self._which, self._channels = ctx, which, channels

    @property
    def shape(self):
        return (self._ctx.width, self._ctx.height)

    def to_numpy(self) -> np.ndarray:
        return self._ctx.download(self._which)

    def from_numpy(self, arr: np.ndarray) -> None:
        self._ctx.upload(self._which, arr)


class PathTracer:
    def __init__(self, config: N.RtpbrConfig, objects, camera: Camera, tonemap: dict | None = None, device: int = 0):
        self.config = config
        self.ctx = N.Context(config, device)
        self.tonemap = tonemap or dict(mode=2, exposure=1.0, gamma=2.2)
        self.image_buffer = Field(self.ctx, N.BUF_IMAGE_BUFFER, 4)     # src/fileds.py:8
        self.image_pixels = Field(self.ctx, N.BUF_IMAGE_PIXELS, 3)     # src/fileds.py:9
        self.ray_buffer = Field(self.ctx, N.BUF_RAY_BUFFER, 10)        # src/fileds.py:7 (family C only)
        self.denoise_pixels = Field(self.ctx, N.BUF_DENOISE_PIXELS, 3) # denoise_test_1.py:53 (after the first denoise())
        self.set_scene(objects)
        self.set_camera(camera)
        if self.tonemap.get("frame"):                 # presets with an animation frame (scenes.bunny_glass(frame=...))
            self.set_frame(int(self.tonemap["frame"]))

    # scene / camera ----------------------------------------------------------------------
    def set_scene(self, objects) -> None:
        self.objects = list(objects)
        self.ctx.set_scene([o.to_native() if isinstance(o, SDFObject) else o for o in self.objects])

    def set_camera(self, camera: Camera) -> None:
        self.camera = camera
        self.ctx.set_camera(camera.to_native() if isinstance(camera, Camera) else camera)

    def set_envmap(self, table: np.ndarray) -> None:
        """Processed environment table, (w, h, 3) f32 (see raytracingpbr_b200.ibl)."""
        self.ctx.set_envmap(table)

    def set_frame(self, frame: int) -> None:
        """`u_frame[None] = frame` (bunny_sdf_glass.py:409): the neural bunny's programmatic rotation + bob."""
        self.ctx.set_frame(frame)

    # frame loop --------------------------------------------------------------------------
    def refresh(self) -> None:
        """kernel refresh(), src/renderer.py:12-22."""
        self.ctx.refresh()

    def pathtrace(self, spp: int = 1) -> None:
        """`spp` x kernel pathtrace()/sample()/render() launches of the reference, in one launch."""
        self.ctx.pathtrace(spp)

    def post_process(self) -> None:
        """kernel post_process(), src/postprocessor.py:24-38 (variant chosen by the preset)."""
        self.ctx.post_process(self.tonemap["mode"], self.tonemap["exposure"], self.tonemap["gamma"])

    def denoise(self, threshold: float = 0.1) -> None:
        """kernel denoise(image_pixels, denoise_pixels, threshold), examples/denoise/denoise_test_1.py:86-118, as a
        deterministic double-buffered pass over the tone-mapped pixels; result in `denoise_pixels`."""
        self.ctx.denoise(threshold)

    def render(self, spp: int = 1, refreshing: bool = False) -> None:
        """render(refreshing), src/renderer.py:25-32."""
        if refreshing:
            self.refresh()
        self.pathtrace(spp)
        self.post_process()

    def sync(self) -> None:
        self.ctx.sync()

    # one process per GPU (torchrun / MPI launches): this process renders one shard -----------------
    def set_shard(self, rank: int, nranks: int, band: int = 4) -> None:
        """This tracer renders the columns i with (i / band) mod nranks == rank (interleaved bands balance the scene's
        cost across ranks; the Philox stream is keyed by the global pixel, so the union of the shards is the
        single-GPU image bit for bit)."""
        self.ctx.set_shard(rank, nranks, band)

    def nccl_init(self, unique_id: bytes, rank: int, nranks: int) -> None:
        """Join the communicator whose 128-byte id rank 0 made with raytracingpbr_b200.nccl_unique_id() and handed to
        the other processes by whatever launched them."""
        self.ctx.nccl_init(unique_id, rank, nranks)

    def reduce_tiles(self, root: int = 0) -> None:
        """Sum of the per-rank sample sums onto `root` (all ranks when root < 0): the only collective, at tonemap time."""
        self.ctx.reduce_tiles(root)

    def close(self) -> None:
        self.ctx.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def imwrite(pixels: np.ndarray, path: str) -> None:
    """ti.tools.imwrite(field, path) stand-in (src/main.py:55): (W,H,3) f32 in [0,1], origin
    bottom-left -> PNG with origin top-left."""
    from PIL import Image
    a = np.clip(np.asarray(pixels), 0.0, 1.0)
    img = (a.transpose(1, 0, 2)[::-1] * 255.0).astype(np.uint8)
    Image.fromarray(img).save(path)


class MultiPathTracer:
    """The PathTracer surface over several GPUs of one box, driven by ONE process (no torch, no launcher): image columns
    are dealt to the GPUs in interleaved bands, every GPU traces its own pixels, and post_process() sums the per-GPU sample
    sums onto GPU 0 with NCCL and tone-maps there (include/rtpbr.h rtpbr_multi_*)."""

    def __init__(self, config: N.RtpbrConfig, objects, camera: Camera, tonemap: dict | None = None, devices=(0,), band: int = 4):
        self.config = config
        self.ctx = N.MultiContext(config, devices, band)
        self.tonemap = tonemap or dict(mode=2, exposure=1.0, gamma=2.2)
        self.image_buffer = Field(self.ctx, N.BUF_IMAGE_BUFFER, 4)     # valid after reduce() / post_process()
        self.image_pixels = Field(self.ctx, N.BUF_IMAGE_PIXELS, 3)
        self.set_scene(objects)
        self.set_camera(camera)
        if self.tonemap.get("frame"):
            self.set_frame(int(self.tonemap["frame"]))

    def set_scene(self, objects) -> None:
        self.objects = list(objects)
        self.ctx.set_scene([o.to_native() if isinstance(o, SDFObject) else o for o in self.objects])

    def set_camera(self, camera: Camera) -> None:
        self.camera = camera
        self.ctx.set_camera(camera.to_native() if isinstance(camera, Camera) else camera)

    def set_envmap(self, table: np.ndarray) -> None:
        self.ctx.set_envmap(table)

    def set_frame(self, frame: int) -> None:
        self.ctx.set_frame(frame)

    def refresh(self) -> None:
        self.ctx.refresh()

    def pathtrace(self, spp: int = 1) -> None:
        self.ctx.pathtrace(spp)

    def reduce(self, root: int = 0) -> None:
        self.ctx.reduce(root)

    def post_process(self) -> None:
        self.ctx.post_process(self.tonemap["mode"], self.tonemap["exposure"], self.tonemap["gamma"])

    def render(self, spp: int = 1) -> None:
        """One finished frame: refresh, trace `spp` samples per pixel across the GPUs, reduce, tonemap."""
        self.refresh()
        self.pathtrace(spp)
        self.post_process()

    def sync(self) -> None:
        self.ctx.sync()

    def close(self) -> None:
        self.ctx.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
