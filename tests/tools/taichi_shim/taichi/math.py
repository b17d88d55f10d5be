"""taichi.math stand-in: vec/mat value types and library functions on fp32 scalars.

Library-function contract (these live inside Taichi, not in the reference):
  dot(a,b) = fmaf(a.z,b.z, fmaf(a.y,b.y, a.x*b.x)) (index order, extended to any length)
  M @ v, v @ M, A @ B: one such dot per output element
  length(v) = sqrt(dot(v,v)); normalize(v) = v * (1 / length(v))
  mix(a,b,t) = a*(1-t) + b*t; cross = a.y*b.z - a.z*b.y, ...
  sin / cos / atan2 / asin: the polynomial routines below (same algorithm and constants as
  oracle/oracle.c, written independently in Python); exp / tan / general pow: binary64 libm
  rounded once.
"""
import math as _m

from ._scalar import F, fmaf, powf, r32, sqrtf

pi = _m.pi
e = _m.e
inf = _m.inf
nan = _m.nan

_SWZ = {"x": 0, "y": 1, "z": 2, "w": 3, "r": 0, "g": 1, "b": 2, "a": 3}


def _flat(args):
    out = []
    for a in args:
        if isinstance(a, Vec):
            out.extend(a._v)
        elif isinstance(a, (tuple, list)):
            out.extend(_flat(a))
        else:
            out.append(a)
    return out


class Vec:
    """Immutable-by-convention fp32 vector (value semantics: every operator returns a new Vec)."""
    __slots__ = ("_v",)
    n = None

    def __init__(self, *args):
        v = _flat(args)
        n = type(self).n
        if n is None:
            n = len(v)
        if len(v) == 0:
            v = [0.0] * n
        elif len(v) == 1:
            v = v * n
        if len(v) != n:
            raise ValueError(f"vec{n} from {len(v)} components")
        object.__setattr__(self, "_v", tuple(F(c) for c in v))

    # ---- access
    def __len__(self):
        return len(self._v)

    def __iter__(self):
        return iter(self._v)

    def __getitem__(self, i):
        return self._v[i]

    def __getattr__(self, name):
        try:
            idx = [_SWZ[c] for c in name]
        except KeyError:
            raise AttributeError(name) from None
        if len(idx) == 1:
            return self._v[idx[0]]
        return _vec_of(len(idx))(*[self._v[i] for i in idx])

    def __setattr__(self, name, value):
        if name == "_v":
            object.__setattr__(self, name, value)
            return
        idx = [_SWZ[c] for c in name]
        v = list(self._v)
        vals = [value] if len(idx) == 1 else list(value)
        for i, x in zip(idx, vals):
            v[i] = F(x)
        object.__setattr__(self, "_v", tuple(v))

    def __setitem__(self, i, value):
        v = list(self._v)
        v[i] = F(value)
        object.__setattr__(self, "_v", tuple(v))

    def copy(self):
        return type(self)(*self._v)

    # ---- arithmetic (element-wise, scalars broadcast)
    def _bin(self, o, f):
        if isinstance(o, Vec):
            if len(o) != len(self):
                raise ValueError("vector length mismatch")
            return _vec_of(len(self))(*[f(a, b) for a, b in zip(self._v, o._v)])
        if isinstance(o, (int, float)):
            return _vec_of(len(self))(*[f(a, o) for a in self._v])
        return NotImplemented

    def __add__(self, o):
        return self._bin(o, lambda a, b: a + b)

    __radd__ = __add__

    def __sub__(self, o):
        return self._bin(o, lambda a, b: a - b)

    def __rsub__(self, o):
        return self._bin(o, lambda a, b: b - a)

    def __mul__(self, o):
        return self._bin(o, lambda a, b: a * b)

    __rmul__ = __mul__

    def __truediv__(self, o):
        return self._bin(o, lambda a, b: a / b)

    def __rtruediv__(self, o):
        return self._bin(o, lambda a, b: b / a)

    def __pow__(self, o):
        return self._bin(o, lambda a, b: powf(a, b))

    def __neg__(self):
        return _vec_of(len(self))(*[-a for a in self._v])

    def __abs__(self):
        return _vec_of(len(self))(*[abs(a) for a in self._v])

    def __matmul__(self, o):
        if isinstance(o, Mat):                       # row vector times matrix
            return _vec_of(o.cols)(*[_dot(self._v, [o._m[k][j] for k in range(o.rows)]) for j in range(o.cols)])
        return NotImplemented

    def max(self):
        out = self._v[0]
        for a in self._v[1:]:
            out = a if a > out else out
        return out

    def min(self):
        out = self._v[0]
        for a in self._v[1:]:
            out = a if a < out else out
        return out

    def norm(self):
        return length(self)

    def normalized(self):
        return normalize(self)

    def dot(self, o):
        return dot(self, o)

    def cross(self, o):
        return cross(self, o)

    def to_list(self):
        return [float(a) for a in self._v]

    def __repr__(self):
        return f"vec{len(self)}({', '.join(repr(float(a)) for a in self._v)})"

    @classmethod
    def field(cls, shape=None, **kw):
        from . import _Field
        return _Field(cls, shape)


class vec2(Vec):
    __slots__ = ()
    n = 2


class vec3(Vec):
    __slots__ = ()
    n = 3


class vec4(Vec):
    __slots__ = ()
    n = 4


def _vec_of(n):
    return {2: vec2, 3: vec3, 4: vec4}[n]


def _dot(a, b):
    acc = F(a[0]) * b[0]
    for x, y in zip(a[1:], b[1:]):
        acc = fmaf(x, y, acc)
    return acc


class Mat:
    __slots__ = ("_m", "rows", "cols")
    n = None

    def __init__(self, *args):
        n = type(self).n
        if len(args) == n and all(isinstance(a, Vec) for a in args):      # rows as vectors
            rows = [list(a._v) for a in args]
        else:
            v = _flat(args)
            if len(v) == 0:
                v = [0.0] * (n * n)
            if len(v) != n * n:
                raise ValueError(f"mat{n} from {len(v)} entries")
            rows = [v[r * n:(r + 1) * n] for r in range(n)]               # row-major
        self._m = tuple(tuple(F(c) for c in row) for row in rows)
        self.rows = self.cols = n

    def __getitem__(self, ij):
        if isinstance(ij, tuple):
            return self._m[ij[0]][ij[1]]
        return self._m[ij]

    def copy(self):
        return type(self)(*[c for row in self._m for c in row])

    def __matmul__(self, o):
        if isinstance(o, Vec):
            return _vec_of(self.rows)(*[_dot(self._m[i], o._v) for i in range(self.rows)])
        if isinstance(o, Mat):
            return type(self)(*[_dot(self._m[i], [o._m[k][j] for k in range(o.rows)])
                                for i in range(self.rows) for j in range(o.cols)])
        return NotImplemented

    def to_list(self):
        return [[float(c) for c in row] for row in self._m]

    def __repr__(self):
        return f"mat{self.rows}({self.to_list()})"


class mat2(Mat):
    __slots__ = ()
    n = 2


class mat3(Mat):
    __slots__ = ()
    n = 3


class mat4(Mat):
    __slots__ = ()
    n = 4


# ------------------------------------------------------------------ element-wise helpers
def _map1(f, x):
    if isinstance(x, Vec):
        return _vec_of(len(x))(*[f(a) for a in x._v])
    return f(x)


def _map2(f, x, y):
    if isinstance(x, Vec) and isinstance(y, Vec):
        return _vec_of(len(x))(*[f(a, b) for a, b in zip(x._v, y._v)])
    if isinstance(x, Vec):
        return _vec_of(len(x))(*[f(a, y) for a in x._v])
    if isinstance(y, Vec):
        return _vec_of(len(y))(*[f(x, b) for b in y._v])
    return f(x, y)


def _max2(a, b):
    a, b = F(a), F(b)
    return a if a > b else b


def _min2(a, b):
    a, b = F(a), F(b)
    return a if a < b else b


def max(*args):  # noqa: A001
    out = args[0]
    for a in args[1:]:
        out = _map2(_max2, out, a)
    return out


def min(*args):  # noqa: A001
    out = args[0]
    for a in args[1:]:
        out = _map2(_min2, out, a)
    return out


def clamp(x, lo, hi):
    return min(max(x, lo), hi)


def sqrt(x):
    return _map1(sqrtf, x)


def pow(x, y):  # noqa: A001
    return _map2(powf, x, y)


def exp(x):
    def f(a):
        try:
            return F(_m.exp(r32(float(a))))
        except OverflowError:
            return F(_m.inf)
    return _map1(f, x)


def tan(x):
    return _map1(lambda a: F(_m.tan(r32(float(a)))), x)


def radians(x):
    k = F(_m.pi / 180.0)
    return _map1(lambda a: F(a) * k, x)


def sign(x):
    return _map1(lambda a: F(1.0) if a > 0 else (F(-1.0) if a < 0 else F(0.0)), x)


def mix(a, b, t):
    return a * (1.0 - t) + b * t


def dot(a, b):
    return _dot(a._v, b._v)


def length(a):
    return sqrtf(_dot(a._v, a._v))


def normalize(a):
    return a * (1.0 / length(a))


def distance(a, b):
    return length(a - b)


def cross(a, b):
    return vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x)


def reflect(i, n):
    return i - 2.0 * dot(n, i) * n


# ------------------------------------------------------------------ transcendental contract
def _sincos(x):
    x = F(x)
    j = float(round(float(x * 0.636619746685028076)))      # rintf: round half to even
    nj = F(-j)
    r = fmaf(nj, float.fromhex("0x1.921fb6p+0"), x)
    r = fmaf(nj, float.fromhex("-0x1.777a5cp-25"), r)
    r = fmaf(nj, float.fromhex("-0x1.ee59dap-50"), r)
    q = int(j)
    r2 = r * r
    sp = fmaf(r2, -1.9515295891e-4, 8.3321608736e-3)
    sp = fmaf(sp, r2, -1.6666654611e-1)
    s = fmaf(sp * r2, r, r)
    cp = fmaf(r2, 2.443315711809948e-5, -1.388731625493765e-3)
    cp = fmaf(cp, r2, 4.166664568298827e-2)
    cp = fmaf(cp, r2, -0.5)
    c = fmaf(cp, r2, 1.0)
    if q & 1:
        s, c = c, s
    if q & 2:
        s = -s
    if (q + 1) & 2:
        c = -c
    return s, c


def sin(x):
    return _map1(lambda a: _sincos(a)[0], x)


def cos(x):
    return _map1(lambda a: _sincos(a)[1], x)


def _atan01(a):
    s = a * a
    p = fmaf(s, 0.00282363896258175373077393, -0.0159569028764963150024414)
    for c in (0.0425049886107444763183594, -0.0748900920152664184570312, 0.106347933411598205566406,
              -0.142027363181114196777344, 0.199926957488059997558594, -0.333331018686294555664062):
        p = fmaf(p, s, c)
    return fmaf(p * s, a, a)


def _atan2(y, x):
    y, x = F(y), F(x)
    ax, ay = abs(x), abs(y)
    if ax > ay:
        mx, mn = ax, ay
    else:
        mx, mn = ay, ax
    a = F(0.0) if mx == 0.0 else mn / mx
    r = _atan01(a)
    if ay > ax:
        r = F(float.fromhex("0x1.921fb6p+0")) - r
    if x < 0.0:
        r = F(3.14159274101257324) - r
    return -r if y < 0.0 else r


def atan2(y, x):
    return _map2(_atan2, y, x)


def _asin(x):
    x = F(x)
    x = F(1.0) if x > 1.0 else (F(-1.0) if x < -1.0 else x)
    return _atan2(x, sqrtf((1.0 - x) * (1.0 + x)))


def asin(x):
    return _map1(_asin, x)
