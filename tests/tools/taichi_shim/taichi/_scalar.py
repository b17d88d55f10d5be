"""fp32 scalar with Taichi's evaluation rules.

`F` is a Python float that always holds a binary32-representable value.  Any arithmetic that
touches an F is rounded to binary32 once per operator (computing in binary64 and rounding is
exact for + - * / sqrt: 53 >= 2*24 + 2), which is what Taichi's default_fp=f32 does for run-time
values.  Expressions made only of Python literals stay Python doubles, exactly like Taichi's
Python-scope constant folding (e.g. `0.5773 * 0.005` in src/sdf.py:80); they are rounded when
they first meet an F.  Library functions (dot, matmul, length, sin, ...) live in math.py and
follow the fp32 contract of the repo.
"""
import ctypes
import math
import struct

_pack = struct.Struct("f")
_libm = ctypes.CDLL("libm.so.6")
_libm.fmaf.restype = ctypes.c_float
_libm.fmaf.argtypes = [ctypes.c_float] * 3


def r32(x) -> float:
    """Round a Python number to the nearest binary32 value (returned as a Python float)."""
    try:
        return _pack.unpack(_pack.pack(x))[0]
    except OverflowError:
        return math.copysign(math.inf, x)


class F(float):
    __slots__ = ()

    def __new__(cls, x=0.0):
        return float.__new__(cls, r32(float(x)))

    # every operator: one binary32 rounding
    def __add__(self, o):
        return _mk(float.__add__(self, _num(o))) if _ok(o) else NotImplemented

    __radd__ = __add__

    def __sub__(self, o):
        return _mk(float.__sub__(self, _num(o))) if _ok(o) else NotImplemented

    def __rsub__(self, o):
        return _mk(float.__rsub__(self, _num(o))) if _ok(o) else NotImplemented

    def __mul__(self, o):
        return _mk(float.__mul__(self, _num(o))) if _ok(o) else NotImplemented

    __rmul__ = __mul__

    def __truediv__(self, o):
        if not _ok(o):
            return NotImplemented
        d = _num(o)
        if d == 0.0:
            return _mk(math.copysign(math.inf, self) * math.copysign(1.0, d) if self != 0.0 else math.nan)
        return _mk(float.__truediv__(self, d))

    def __rtruediv__(self, o):
        if not _ok(o):
            return NotImplemented
        n = _num(o)
        if self == 0.0:
            return _mk(math.copysign(math.inf, n) * math.copysign(1.0, self) if n != 0.0 else math.nan)
        return _mk(n / float(self))

    def __neg__(self):
        return _mk(float.__neg__(self))

    def __pos__(self):
        return self

    def __abs__(self):
        return _mk(float.__abs__(self))

    def __pow__(self, e, mod=None):
        return powf(self, e)

    def __rpow__(self, b, mod=None):
        return powf(F(b), self)

    # comparisons: the other side is rounded to binary32 first (Taichi casts literals to f32)
    def __lt__(self, o):
        return float.__lt__(self, _num(o))

    def __le__(self, o):
        return float.__le__(self, _num(o))

    def __gt__(self, o):
        return float.__gt__(self, _num(o))

    def __ge__(self, o):
        return float.__ge__(self, _num(o))

    def __eq__(self, o):
        return float.__eq__(self, _num(o)) if _ok(o) else NotImplemented

    def __ne__(self, o):
        return float.__ne__(self, _num(o)) if _ok(o) else NotImplemented

    __hash__ = float.__hash__

    def __repr__(self):
        return f"F({float.__repr__(self)})"


class I(int):
    """Run-time i32 value (struct members, field elements).  Arithmetic with a float yields an
    F -- Taichi casts the i32 to f32 and computes in f32 -- so expressions such as
    `ray.depth * (1.0 / MAX_RAYTRACE)` (src/pathtracer.py:68) are rounded like the reference."""
    __slots__ = ()

    def _wrap(self, r, o):
        if r is NotImplemented:
            return r
        return I(r) if isinstance(o, int) and not isinstance(o, bool) or isinstance(o, bool) else r

    def __add__(self, o):
        return F(int(self)) + o if isinstance(o, float) else I(int.__add__(self, o))

    __radd__ = __add__

    def __sub__(self, o):
        return F(int(self)) - o if isinstance(o, float) else I(int.__sub__(self, o))

    def __rsub__(self, o):
        return o - F(int(self)) if isinstance(o, float) else I(int.__rsub__(self, o))

    def __mul__(self, o):
        return F(int(self)) * o if isinstance(o, float) else I(int.__mul__(self, o))

    __rmul__ = __mul__

    def __truediv__(self, o):
        return F(int(self)) / o

    def __rtruediv__(self, o):
        return F(o) / F(int(self))

    def __neg__(self):
        return I(int.__neg__(self))


def _ok(o):
    return isinstance(o, (int, float, bool))


def _num(o) -> float:
    """Operand as a binary32-valued Python float."""
    if type(o) is F:
        return float.__float__(o)
    return r32(float(o))


def _mk(x: float) -> F:
    return float.__new__(F, r32(x))


def fmaf(a, b, c) -> F:
    """Correctly rounded binary32 fused multiply-add (libm fmaf through ctypes)."""
    return float.__new__(F, _libm.fmaf(a, b, c))


def sqrtf(x) -> F:
    x = _num(x)
    return _mk(math.sqrt(x)) if x >= 0.0 else _mk(math.nan)


def powf(x, e) -> F:
    """pow contract: exponent 5.0 (Schlick) = ((x*x)*(x*x))*x; anything else = binary64 libm
    pow rounded once (only used by tone mapping / env-map preprocessing, which run on the host)."""
    x = F(x)
    e = _num(e)
    if e == 5.0:
        x2 = x * x
        return (x2 * x2) * x
    if x == 0.0:
        return F(0.0) if e > 0 else F(math.inf)
    if x < 0.0 and e != int(e):
        return F(math.nan)
    try:
        return F(math.pow(x, e))
    except OverflowError:
        return F(math.inf)
