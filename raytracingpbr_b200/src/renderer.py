"""src/renderer.py -- refresh() and render(refreshing)."""
from . import _runtime, config
from .camera import smooth
from .pathtracer import pathtrace
from .postprocessor import post_process


def refresh():
    """src/renderer.py:12-22"""
    _runtime.tracer().refresh()


def render(refreshing):
    """src/renderer.py:25-32"""
    if refreshing or smooth.moving[None]:
        refresh()
    pathtrace(config.SAMPLES_PER_FRAME)
    post_process()
