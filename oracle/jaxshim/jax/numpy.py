import numpy as _np

from ._core import Arr, asarr

newaxis = None
bool = _np.bool_
bool_ = _np.bool_
ndarray = Arr


class _DT:
    """jnp.int32 & co: callable like a constructor, usable as a dtype."""

    def __init__(self, dt):
        self.dtype = _np.dtype(dt)

    def __call__(self, x):
        return asarr(_np.asarray(x).astype(self.dtype), self.dtype)

    def __eq__(self, other):
        try:
            return _np.dtype(other) == self.dtype
        except TypeError:
            return False

    def __hash__(self):
        return hash(self.dtype)


def _dt(d):
    if isinstance(d, _DT):
        return d.dtype
    return d


int32, uint32, uint8, float32, uint64, float64, int8 = (_DT(d) for d in ("int32", "uint32", "uint8", "float32", "uint64", "float64", "int8"))
bool = _DT("bool")
bool_ = bool
float_ = float32


def _wrap(f):
    def g(*a, **k):
        if "dtype" in k:
            k["dtype"] = _dt(k["dtype"])
        with _np.errstate(over="ignore"):
            r = f(*a, **k)
        if isinstance(r, _np.ndarray) or _np.isscalar(r):
            r = asarr(r)
        return r

    return g


def zeros(shape, dtype=float32):
    return asarr(_np.zeros(shape, _dt(dtype)))


def ones(shape, dtype=float32):
    return asarr(_np.ones(shape, _dt(dtype)))


def array(x, dtype=None):
    if dtype is None:
        return asarr(_np.array(x))
    return asarr(_np.array(x, dtype=_dt(dtype)), _dt(dtype))


def asarray(x, dtype=None):
    return array(x, dtype)


def arange(*a, dtype=None):
    return asarr(_np.arange(*a, dtype=_dt(dtype) if dtype is not None else _np.int32))


def astype_patch():
    pass


for _n in ("where", "clip", "zeros_like", "ones_like", "minimum", "maximum", "stack", "concatenate", "roll", "atleast_1d",
           "reshape", "swapaxes", "tile", "max", "min", "sum", "tanh", "sqrt", "exp", "matmul", "dot", "square", "abs",
           "argmax", "linspace", "full", "split", "squeeze", "any", "all", "mean", "log", "transpose", "take_along_axis",
           "expand_dims", "cumsum", "power", "isnan", "logical_and", "logical_or", "logical_not"):
    globals()[_n] = _wrap(getattr(_np, _n))


def unpackbits(a, axis=None, count=None, bitorder="big"):
    return asarr(_np.unpackbits(_np.asarray(a), axis=axis, count=count, bitorder=bitorder))


class _Finfo:
    def __init__(self, dt):
        fi = _np.finfo(_dt(dt))
        self.min, self.max, self.tiny, self.eps = fi.dtype.type(fi.min), fi.dtype.type(fi.max), fi.dtype.type(fi.tiny), fi.dtype.type(fi.eps)


def finfo(dt):
    return _Finfo(dt)


# ndarray.astype(jnp.float32) must accept the _DT wrappers
_orig_astype = _np.ndarray.astype


def _astype(self, dtype, *a, **k):
    return asarr(_orig_astype(_np.asarray(self), _dt(dtype), *a, **k))


Arr.astype = _astype
_orig_clip = _np.ndarray.clip


def _clip(self, min=None, max=None, **k):
    return asarr(_np.clip(_np.asarray(self), min, max))


Arr.clip = _clip
