"""src/ibl.py -- environment image.  `sky_color` / `Image.texture` run inside the CUDA kernel."""
from .. import ibl as _ibl
from . import _runtime
from .camera import camera_gamma


class Image:
    """src/ibl.py:12-29"""
    def __init__(self, path: str):
        self.img = _ibl.imread(path)                 # uint8 (W, H, 3), like ti.tools.imread
        self.table = None

    def process(self, exposure: float, gamma: float):
        self.table = _ibl.process(self.img, exposure, gamma)
        _runtime.set_env(self.table)


hdr_map = None


def load(path: str = "assets/Tokyo_BigSight_3k.hdr", exposure: float = 1.4, gamma: float = camera_gamma):
    """src/ibl.py:32-33 (done at import time by the reference; explicit here)."""
    global hdr_map
    hdr_map = Image(path)
    hdr_map.process(exposure=exposure, gamma=gamma)
    return hdr_map
