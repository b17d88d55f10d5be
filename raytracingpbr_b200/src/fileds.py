"""src/fileds.py -- the global fields, as read-only views of the device buffers."""
from .. import _native as N
from . import _runtime


class _FieldView:
    def __init__(self, which):
        self._which = which

    @property
    def shape(self):
        from . import config
        return tuple(config.image_resolution)

    def to_numpy(self):
        return _runtime.tracer().ctx.download(self._which)

    def from_numpy(self, arr):
        _runtime.tracer().ctx.upload(self._which, arr)


ray_buffer = _FieldView(N.BUF_RAY_BUFFER)       # src/fileds.py:7
image_buffer = _FieldView(N.BUF_IMAGE_BUFFER)   # :8
image_pixels = _FieldView(N.BUF_IMAGE_PIXELS)   # :9
u_frame = _runtime.ScalarField(0, camera=False)  # :15
diff_buffer = _FieldView(N.BUF_DIFF_BUFFER)     # :21 (ADAPTIVE_SAMPLING only)
diff_pixels = _FieldView(N.BUF_DIFF_PIXELS)     # :22
