"""src/pathtracer.py:94-103 -- kernel pathtrace()."""
from . import _runtime


def pathtrace(launches: int = 1):
    """One call = `launches` reference launches of kernel pathtrace() (replayed in one CUDA launch)."""
    _runtime.tracer().pathtrace(launches)
