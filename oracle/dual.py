"""
Forward-mode dual numbers on numpy arrays (oracle, test infrastructure).

Stands in for pycppad's ``a_float`` + ``adfun.jacobian`` (reference
devices/cppaddev.py:113-157): a ``Dual`` carries a value array ``v`` of shape
``(n,)`` and a tangent array ``d`` of shape ``(P, n)``; every operation
propagates exact first derivatives.  ``condassign``/``safe_exp`` follow
cppaddev.py:56-96: strict ``b > 0`` selection applied to value *and* tangents
(the un-selected branch may be NaN/inf and must not leak), ``d|x|/dx = 0`` at
0 and NaN comparisons false, as CppAD does.

``OPS`` counts scalar FP64 operations of the primal + P tangent lanes with the
weights of SURVEY section 8(d) (add/sub/mul/neg/abs/select = 1, div = 1,
sqrt = 1, exp/log/tan/tanh/pow = 1 and tallied separately as ``trans``); it
defines the algorithmic FLOP count used by bench.py's roofline.
"""
import numpy as np

OPS = dict(flop=0, trans=0)


def reset_ops():
    OPS['flop'] = 0
    OPS['trans'] = 0


def _count(flop, trans=0):
    OPS['flop'] += flop
    OPS['trans'] += trans


class Dual(object):
    __slots__ = ('v', 'd')
    __array_priority__ = 1000.

    def __init__(self, v, d):
        self.v = v
        self.d = d

    @property
    def P(self):
        return self.d.shape[0]

    # ---- arithmetic ---------------------------------------------------
    def __neg__(self):
        _count(1 + self.P)
        return Dual(-self.v, -self.d)

    def __pos__(self):
        return self

    def __add__(self, o):
        if isinstance(o, Dual):
            _count(1 + self.P)
            return Dual(self.v + o.v, self.d + o.d)
        _count(1)
        return Dual(self.v + o, self.d)
    __radd__ = __add__

    def __sub__(self, o):
        if isinstance(o, Dual):
            _count(1 + self.P)
            return Dual(self.v - o.v, self.d - o.d)
        _count(1)
        return Dual(self.v - o, self.d)

    def __rsub__(self, o):
        _count(1 + self.P)
        return Dual(o - self.v, -self.d)

    def __mul__(self, o):
        if isinstance(o, Dual):
            _count(1 + 3 * self.P)
            return Dual(self.v * o.v, self.d * o.v + self.v * o.d)
        _count(1 + self.P)
        return Dual(self.v * o, self.d * o)
    __rmul__ = __mul__

    def __truediv__(self, o):
        if isinstance(o, Dual):
            _count(1 + 3 * self.P)
            z = self.v / o.v
            return Dual(z, (self.d - z * o.d) / o.v)
        _count(1 + self.P)
        return Dual(self.v / o, self.d / o)

    def __rtruediv__(self, o):
        _count(2 + 2 * self.P)
        z = o / self.v
        return Dual(z, (-z / self.v) * self.d)

    def __pow__(self, p):
        if isinstance(p, Dual):
            return exp(p * log(self))
        if p == 2:
            return self * self
        if p == 3:
            return self * self * self
        return power(self, p)

    def __rpow__(self, base):
        return exp(self * np.log(base))

    def __abs__(self):
        return dabs(self)

    # numpy ufuncs applied to a dual (the reference's device code calls
    # np.exp / np.sqrt / np.abs ... on pycppad a_float values, which numpy
    # forwards to the object's methods; here they map to the functions below)
    def __array_ufunc__(self, ufunc, method, *args, **kwargs):
        out = kwargs.pop('out', None)          # in-place: vec *= dual
        if method != '__call__' or kwargs:
            return NotImplemented
        if out is not None:
            res = self.__array_ufunc__(ufunc, method, *args)
            if res is NotImplemented or not _is_objarr(out[0]):
                return NotImplemented
            out[0][...] = res
            return out[0]
        f = _UFUNCS.get(ufunc.__name__)
        if f is None:
            return NotImplemented
        if any(_is_objarr(a) for a in args):
            # dual (op) object array of duals: element by element, as numpy
            # does for arrays of pycppad a_float
            shape = [a for a in args if _is_objarr(a)][0].shape
            out = np.empty(shape, dtype=object)
            for i in np.ndindex(*shape):
                out[i] = f(*[a[i] if _is_objarr(a) else a for a in args])
            return out
        return f(*args)

    # comparisons act on values (used only through condassign in models)
    def __gt__(self, o):
        return self.v > _val(o)

    def __lt__(self, o):
        return self.v < _val(o)

    def __ge__(self, o):
        return self.v >= _val(o)

    def __le__(self, o):
        return self.v <= _val(o)


def _val(x):
    return x.v if isinstance(x, Dual) else x


def _is_objarr(x):
    return isinstance(x, np.ndarray) and x.dtype == object


def _defer(method):
    """Binary operators hand object arrays (vectors of duals) back to numpy,
    which then applies the operation element by element."""
    def wrapped(self, o):
        if _is_objarr(o):
            return NotImplemented
        return method(self, o)
    wrapped.__name__ = method.__name__
    return wrapped


for _name in ('__add__', '__radd__', '__sub__', '__rsub__', '__mul__',
              '__rmul__', '__truediv__', '__rtruediv__'):
    setattr(Dual, _name, _defer(getattr(Dual, _name)))


def seed(x):
    """Independent variables: x of shape (P, n) -> list of P Duals."""
    x = np.asarray(x, dtype=float)
    P, n = x.shape
    out = []
    for k in range(P):
        d = np.zeros((P, n))
        d[k] = 1.
        out.append(Dual(x[k].copy(), d))
    return out


def _unary(x, f, df, trans):
    """z = f(x); dz = df(x, z) * dx"""
    if isinstance(x, Dual):
        z = f(x.v)
        _count(2 + x.P, 1 if trans else 0)
        return Dual(z, df(x.v, z) * x.d)
    return f(x)


def exp(x):
    with np.errstate(over='ignore', invalid='ignore'):
        return _unary(x, np.exp, lambda v, z: z, True)


def log(x):
    with np.errstate(divide='ignore', invalid='ignore'):
        return _unary(x, np.log, lambda v, z: 1. / v, True)


def sqrt(x):
    with np.errstate(divide='ignore', invalid='ignore'):
        return _unary(x, np.sqrt, lambda v, z: .5 / z, False)


def tan(x):
    with np.errstate(invalid='ignore', over='ignore'):
        return _unary(x, np.tan, lambda v, z: 1. + z * z, True)


def tanh(x):
    return _unary(x, np.tanh, lambda v, z: 1. - z * z, True)


def cosh(x):
    with np.errstate(over='ignore'):
        return _unary(x, np.cosh, lambda v, z: np.sinh(v), True)


def sin(x):
    return _unary(x, np.sin, lambda v, z: np.cos(v), True)


def cos(x):
    return _unary(x, np.cos, lambda v, z: -np.sin(v), True)


def dabs(x):
    # CppAD: derivative of abs is sign(x), with sign(0) = 0
    return _unary(x, np.abs, lambda v, z: np.sign(v), False)


def power(x, p):
    """x**p for a constant exponent p."""
    with np.errstate(invalid='ignore', divide='ignore', over='ignore'):
        if isinstance(x, Dual):
            z = np.power(x.v, p)
            _count(3 + x.P, 1)
            return Dual(z, (p * z / x.v) * x.d)
        return np.power(x, p)


_UFUNCS = dict(
    exp=exp, log=log, sqrt=sqrt, tan=tan, tanh=tanh, cosh=cosh, sin=sin,
    cos=cos, absolute=dabs, fabs=dabs, negative=lambda a: -a,
    add=lambda a, b: a + b if isinstance(a, Dual) else b.__radd__(a),
    subtract=lambda a, b: a - b if isinstance(a, Dual) else b.__rsub__(a),
    multiply=lambda a, b: a * b if isinstance(a, Dual) else b.__rmul__(a),
    true_divide=lambda a, b: a / b if isinstance(a, Dual)
    else b.__rtruediv__(a),
    divide=lambda a, b: a / b if isinstance(a, Dual) else b.__rtruediv__(a),
    power=lambda a, b: a ** b if isinstance(a, Dual) else b.__rpow__(a),
    greater=lambda a, b: _val(a) > _val(b),
    less=lambda a, b: _val(a) < _val(b),
    greater_equal=lambda a, b: _val(a) >= _val(b),
    less_equal=lambda a, b: _val(a) <= _val(b))


def condassign(b, c, d):
    """(b > 0) ? c : d on values and tangents (cppaddev.py:72-96)."""
    mask = _val(b) > 0.
    if isinstance(c, Dual) or isinstance(d, Dual):
        ref = c if isinstance(c, Dual) else d
        P = ref.P
        cv, cd = (c.v, c.d) if isinstance(c, Dual) else (c, 0.)
        dv, dd = (d.v, d.d) if isinstance(d, Dual) else (d, 0.)
        _count(1 + P)
        v = np.where(mask, cv, dv)
        t = np.where(mask, cd, dd)
        if np.ndim(t) < 2:
            t = np.broadcast_to(t, ref.d.shape).copy()
        return Dual(v, t)
    return np.where(mask, c, d)


def safe_exp(x):
    """exp(x) below 50, linear continuation above (cppaddev.py:56-70)."""
    threshold = 50.
    c = exp(x)
    d = np.exp(threshold) * (x - threshold + 1.)
    # condexp_lt(x, threshold, c, d)
    return condassign(threshold - x, c, d)


def value(x):
    return _val(x)


def jac_rows(outs, P, n):
    """Stack outputs into (values[nout, n], jac[nout, P, n])."""
    vals = np.empty((len(outs), n))
    jac = np.zeros((len(outs), P, n))
    for r, o in enumerate(outs):
        if isinstance(o, Dual):
            vals[r] = o.v
            jac[r] = o.d
        else:
            vals[r] = o
    return vals, jac
