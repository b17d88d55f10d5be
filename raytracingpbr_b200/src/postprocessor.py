"""src/postprocessor.py:24-38 -- kernel post_process() (average -> exposure -> gamma -> ACES -> clamp)."""
from . import _runtime


def post_process():
    _runtime.tracer().post_process()
