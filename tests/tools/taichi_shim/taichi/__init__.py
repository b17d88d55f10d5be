"""Minimal scalar-interpreter stand-in for the `taichi` package -- TEST TOOLING ONLY.

Purpose: execute the reference's own Python source files (HK-SHAO/RayTracingPBR) without
Taichi, one pixel at a time, in fp32, so that tests/tools/gen_golden.py can produce golden
vectors from the reference code itself.  Kernels run as plain Python loops; `ti.random()` is
served by a hook the generator installs (the Philox stream contract of DESIGN.md section 4);
struct / vector values have value semantics like Taichi locals.  Nothing in the product or the
oracle imports this package.
"""
import copy as _copy
import functools as _ft
import inspect as _inspect
import itertools as _it

import numpy as _np

from . import math  # noqa: F401
from ._scalar import F, I, r32
from .math import Mat, Vec, mat2, mat3, mat4, vec2, vec3, vec4

f32 = float
f64 = float
i32 = int
u8 = int
gpu = "gpu"
cpu = "cpu"
cuda = "cuda"
ij = "ij"
i = "i"


def init(*a, **kw):
    return None


# ------------------------------------------------------------------------------ RNG hook
class _Rng:
    """gen_golden.py installs `next` (a callable returning the next uniform of the current
    pixel's stream) and reads `pixel` / bumps `launch`."""
    def __init__(self):
        self.launch = 0
        self.pixel = 0
        self.n = 0
        self.next = None


rng = _Rng()
pixel_filter = None     # optional callable (i, j) -> bool restricting 2-D struct-for loops (see _Field.__iter__)


def random(dtype=float):
    if rng.next is None:
        raise RuntimeError("taichi shim: no RNG hook installed")
    v = rng.next(rng.pixel, rng.launch, rng.n)
    rng.n += 1
    return F(v)


# ------------------------------------------------------------------------------ values
def _cast(tp, v):
    """Convert `v` to the declared field type (Taichi casts on store)."""
    if tp is float or tp is f32:
        return F(v)
    if tp is int:
        return I(int(v))
    if tp is bool:
        return bool(v)
    if isinstance(tp, type) and issubclass(tp, Vec):
        return tp(v) if not isinstance(v, Vec) else tp(*v._v)
    if isinstance(tp, type) and issubclass(tp, Mat):
        return v.copy() if isinstance(v, Mat) else tp(v)
    if isinstance(tp, type) and issubclass(tp, _Struct):
        return v._value_copy() if isinstance(v, _Struct) else tp()
    return v


def _zero(tp):
    if tp is float or tp is f32:
        return F(0.0)
    if tp is int:
        return I(0)
    if tp is bool:
        return False
    return tp()


class _Struct:
    """Base of @ti.dataclass / ti.types.struct types.  Value semantics: copies are made when a
    struct is read from a field, passed to a ti.func or stored into another struct.  A copy read
    from a field keeps a back-reference so that `field[i].a.b = x` writes through (Taichi's
    store-to-field), while reads always see the copy's own snapshot."""
    _fields = {}

    def __init__(self, *args, **kw):
        object.__setattr__(self, "_origin", None)
        names = list(self._fields)
        vals = dict(zip(names, args))
        vals.update(kw)
        for name, tp in self._fields.items():
            object.__setattr__(self, name, _cast(tp, vals[name]) if name in vals else _zero(tp))

    def __setattr__(self, name, value):
        tp = self._fields[name]
        v = _cast(tp, value)
        object.__setattr__(self, name, v)
        org = self._origin
        if org is not None:
            object.__setattr__(org, name, _cast(tp, value))

    def _value_copy(self, origin=None):
        c = object.__new__(type(self))
        object.__setattr__(c, "_origin", origin)
        for name, tp in self._fields.items():
            v = getattr(self, name)
            if isinstance(v, _Struct):
                v = v._value_copy(origin=v if origin is not None else None)
            elif isinstance(v, Mat):
                v = v.copy()
            object.__setattr__(c, name, v)
        return c

    @classmethod
    def field(cls, shape=None, **kw):
        return _Field(cls, shape)

    def __repr__(self):
        return f"{type(self).__name__}({', '.join(f'{k}={getattr(self, k)!r}' for k in self._fields)})"


def _make_struct(name, fields, methods=None):
    ns = {"_fields": dict(fields)}
    if methods:
        ns.update(methods)
    return type(name, (_Struct,), ns)


def dataclass(cls):
    fields = dict(getattr(cls, "__annotations__", {}))
    methods = {k: v for k, v in vars(cls).items() if callable(v) and not k.startswith("__")}
    return _make_struct(cls.__name__, fields, methods)


class _Types:
    @staticmethod
    def struct(**fields):
        return _make_struct("struct", fields)

    @staticmethod
    def vector(n, dtype=float):
        return {2: vec2, 3: vec3, 4: vec4}[n]


types = _Types()


# ------------------------------------------------------------------------------ fields
class _Field:
    def __init__(self, tp, shape=None):
        self.tp = tp
        self.shape = None
        self._data = None
        if shape is not None:
            self._alloc(shape)

    def _alloc(self, shape):
        if isinstance(shape, int):
            shape = (shape,)
        self.shape = tuple(int(s) for s in shape)
        n = 1
        for s in self.shape:
            n *= s
        self._data = [_zero(self.tp) for _ in range(n)]

    def _idx(self, key):
        if key is None or key == ():
            return 0
        if not isinstance(key, tuple):
            key = (key,)
        k = 0
        for a, s in zip(key, self.shape):
            a = int(a)
            if not 0 <= a < s:
                raise IndexError(f"field index {key} out of bounds for shape {self.shape}")
            k = k * s + a
        return k

    def __getitem__(self, key):
        v = self._data[self._idx(key)]
        if isinstance(v, _Struct):
            return v._value_copy(origin=v)
        return v

    def __setitem__(self, key, value):
        self._data[self._idx(key)] = _cast(self.tp, value)

    def __iter__(self):
        """Struct-for: yields indices; a 2-D iteration is the pixel loop of a sampling kernel, so
        it also positions the RNG hook on the pixel's stream (pixel = i * H + j, n = 0)."""
        if len(self.shape) == 1:
            for a in range(self.shape[0]):
                yield a
            return
        for key in _it.product(*[range(s) for s in self.shape]):
            if len(self.shape) == 2:
                if pixel_filter is not None and not pixel_filter(key[0], key[1]):
                    continue       # gen_golden.py renders a subset of the pixels of a large image (pixels are independent)
                rng.pixel = key[0] * self.shape[1] + key[1]
                rng.n = 0
            yield key

    def fill(self, v):
        for k in range(len(self._data)):
            self._data[k] = _cast(self.tp, v)

    def from_numpy(self, arr):
        arr = _np.asarray(arr)
        if self.shape is None:
            self._alloc(arr.shape[:arr.ndim - (0 if self.tp in (float, int) else 1)])
        flat = arr.reshape(len(self._data), -1)
        for k in range(len(self._data)):
            self._data[k] = _cast(self.tp, flat[k].tolist() if flat.shape[1] > 1 else flat[k, 0])

    def to_numpy(self):
        if isinstance(self._data[0], Vec):
            a = _np.array([v.to_list() for v in self._data], dtype=_np.float32)
            return a.reshape(self.shape + (a.shape[1],))
        if isinstance(self._data[0], _Struct):
            raise TypeError("to_numpy of a struct field: read members instead")
        return _np.array(self._data, dtype=_np.float32 if self.tp is float else _np.int32).reshape(self.shape)

    def member(self, name):
        """numpy array of one (possibly nested, dotted) member of a struct field."""
        out = []
        for v in self._data:
            for part in name.split("."):
                v = getattr(v, part)
            out.append(v.to_list() if isinstance(v, Vec) else v)
        a = _np.array(out)
        return a.reshape(self.shape + a.shape[1:])


def field(dtype=float, shape=None, **kw):
    return _Field(dtype, shape)


class _VectorNS:
    @staticmethod
    def field(n, dtype=float, shape=None, **kw):
        return _Field({2: vec2, 3: vec3, 4: vec4}[n], shape)


Vector = _VectorNS()


class _SNode:
    def __init__(self, shape):
        self.shape = shape

    def place(self, *fields):
        for f in fields:
            f._alloc(self.shape)

    def dense(self, axes, shape):
        return _SNode(shape)


class _Root:
    def dense(self, axes, shape):
        return _SNode(shape)


root = _Root()


# ------------------------------------------------------------------------------ decorators
def static(x, *rest):
    return x if not rest else (x,) + rest


def template():
    return "template"


def _by_value(v):
    return v._value_copy() if isinstance(v, _Struct) else v


_BUILTIN_OVERRIDES = {"max": math.max, "min": math.min}


def func(fn):
    """ti.func: arguments are passed by value; Python builtins max/min act element-wise."""
    for k, v in _BUILTIN_OVERRIDES.items():
        fn.__globals__.setdefault(k, v)

    @_ft.wraps(fn)
    def wrapper(*args, **kw):
        return fn(*[_by_value(a) for a in args], **{k: _by_value(v) for k, v in kw.items()})
    wrapper._ti_func = True
    return wrapper


def kernel(fn):
    for k, v in _BUILTIN_OVERRIDES.items():
        fn.__globals__.setdefault(k, v)
    sig = _inspect.signature(fn)

    @_ft.wraps(fn)
    def wrapper(*args, **kw):
        bound = sig.bind(*args, **kw)
        conv = {}
        for name, val in bound.arguments.items():
            ann = sig.parameters[name].annotation
            conv[name] = _cast(ann, val) if ann is not _inspect.Parameter.empty and name != "self" else val
        return fn(**conv)
    return wrapper


def data_oriented(cls):
    return cls


# ------------------------------------------------------------------------------ tools / ui stubs
class _Tools:
    imread = None      # installed by the generator (synthetic environment image)

    @staticmethod
    def imwrite(*a, **kw):
        return None


tools = _Tools()

from . import ui  # noqa: E402,F401
