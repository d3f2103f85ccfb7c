"""Mini-haiku: just enough of dm-haiku for network/fully_connected.py + network/hashes.py
(module name scoping `fc_az_net/linear_3`, get_parameter / get_state / set_state, transform_with_state)."""
import types

import numpy as _np

from jax._core import asarr

TransformedWithState = object
MutableParams = dict
MutableState = dict
Params = dict
State = dict


class _Frame:
    def __init__(self, params, state, rng, initialising):
        self.params, self.state, self.initialising = params, state, initialising
        self.rng = _np.random.default_rng(rng)
        self.scope = []
        self.counters = {}


_frames = []


def _frame():
    assert _frames, "haiku shim: modules must be used inside transform_with_state"
    return _frames[-1]


def _wrap_method(fn):
    def wrapped(self, *a, **k):
        fr = _frame()
        fr.scope.append(self.module_name)
        try:
            return fn(self, *a, **k)
        finally:
            fr.scope.pop()

    wrapped.__wrapped__ = fn
    return wrapped


class _ModuleMeta(type):
    def __new__(mcs, name, bases, ns):
        for k, v in list(ns.items()):
            if isinstance(v, types.FunctionType) and (k == "__call__" or not k.startswith("__")):
                ns[k] = _wrap_method(getattr(v, "__wrapped__", v))
        return super().__new__(mcs, name, bases, ns)


class Module(metaclass=_ModuleMeta):
    def __init__(self, name=None):
        fr = _frame()
        base = name or _snake(type(self).__name__)
        prefix = "/".join(fr.scope[-1:])  # nested under the module whose method is executing
        key = (prefix, base)
        n = fr.counters.get(key, 0)
        fr.counters[key] = n + 1
        local = base if n == 0 else f"{base}_{n}"
        self.module_name = f"{prefix}/{local}" if prefix else local


def _snake(s):
    out = ""
    for i, c in enumerate(s):
        if c.isupper() and i and not s[i - 1].isupper():
            out += "_"
        out += c.lower()
    return out


def get_parameter(name, shape, dtype=_np.float32, init=None):
    fr = _frame()
    mod = fr.scope[-1]
    bucket = fr.params.setdefault(mod, {})
    if name not in bucket:
        assert fr.initialising, f"missing parameter {mod}/{name}"
        bucket[name] = asarr(init(shape, dtype, fr.rng))
    assert tuple(bucket[name].shape) == tuple(shape), (mod, name, bucket[name].shape, shape)
    return bucket[name]


def get_state(name, shape=None, dtype=_np.float32, init=None):
    fr = _frame()
    mod = fr.scope[-1]
    bucket = fr.state.setdefault(mod, {})
    if name not in bucket:
        assert fr.initialising, f"missing state {mod}/{name}"
        bucket[name] = asarr(init(shape, getattr(dtype, "dtype", dtype), fr.rng))
    return bucket[name]


def set_state(name, value):
    fr = _frame()
    fr.state.setdefault(fr.scope[-1], {})[name] = value


class initializers:
    class Constant:
        def __init__(self, c):
            self.c = c

        def __call__(self, shape, dtype, rng):
            return _np.full(shape, self.c, dtype=dtype)

    class RandomNormal:
        def __init__(self, stddev=1.0, mean=0.0):
            self.s, self.m = stddev, mean

        def __call__(self, shape, dtype, rng):
            return (rng.standard_normal(shape) * self.s + self.m).astype(dtype)

    class TruncatedNormal:
        def __init__(self, stddev=1.0, mean=0.0):
            self.s, self.m = stddev, mean

        def __call__(self, shape, dtype, rng):
            return (rng.standard_normal(shape).clip(-2, 2) * self.s + self.m).astype(dtype)


class Linear(Module):
    def __init__(self, output_size, name=None):
        super().__init__(name=name or "linear")
        self.output_size = output_size

    def __call__(self, x):
        x = _np.asarray(x)
        i = x.shape[-1]
        w = get_parameter("w", [i, self.output_size], init=initializers.TruncatedNormal(1.0 / _np.sqrt(i)))
        b = get_parameter("b", [self.output_size], init=initializers.Constant(0.0))
        return asarr(_np.dot(x, _np.asarray(w)) + _np.asarray(b))


class Conv2D(Module):
    """hk.Conv2D(output_channels, kernel_shape): NHWC input, HWIO weights, stride 1, SAME zero padding, with bias."""

    def __init__(self, output_channels, kernel_shape, name=None):
        super().__init__(name=name or "conv2_d")
        self.output_channels, self.k = output_channels, int(kernel_shape)

    def __call__(self, x):
        x = _np.asarray(x, _np.float32)
        B, H, W, Ci = x.shape
        k, Co = self.k, self.output_channels
        fan_in = k * k * Ci
        w = _np.asarray(get_parameter("w", [k, k, Ci, Co], init=initializers.TruncatedNormal(1.0 / _np.sqrt(fan_in))))
        b = _np.asarray(get_parameter("b", [Co], init=initializers.Constant(0.0)))
        p = k // 2
        xp = _np.pad(x, ((0, 0), (p, p), (p, p), (0, 0)))
        y = _np.zeros((B, H, W, Co), _np.float32)
        for kh in range(k):
            for kw in range(k):
                y += _np.tensordot(xp[:, kh:kh + H, kw:kw + W, :], w[kh, kw], axes=([3], [0])).astype(_np.float32)
        return asarr(y + b)


class ExponentialMovingAverage(Module):
    def __init__(self, decay, name=None):
        super().__init__(name=name or "exponential_moving_average")
        self.decay = decay

    def average(self, shape):
        return get_state("average", shape, dtype=_np.dtype(_np.float32), init=initializers.Constant(0.0))


class BatchNorm(Module):
    """hk.BatchNorm(create_scale, create_offset, decay_rate) -- inference only (is_training=False, test_local_stats=False):
    (x - mean_ema.average) * (scale * rsqrt(var_ema.average + eps)) + offset, eps = 1e-5, statistics over all but the channel axis."""

    def __init__(self, create_scale, create_offset, decay_rate, eps=1e-5, name=None):
        super().__init__(name=name or "batch_norm")
        self.eps = eps
        fr = _frame()
        fr.scope.append(self.module_name)
        try:  # haiku creates the two moving averages as children named ~/mean_ema, ~/var_ema
            self.mean_ema = ExponentialMovingAverage(decay_rate, name="~/mean_ema")
            self.var_ema = ExponentialMovingAverage(decay_rate, name="~/var_ema")
        finally:
            fr.scope.pop()

    def __call__(self, x, is_training, test_local_stats=False):
        assert not is_training and not test_local_stats, "haiku shim: BatchNorm is inference-only"
        x = _np.asarray(x, _np.float32)
        c = x.shape[-1]
        shape = [1] * (x.ndim - 1) + [c]
        scale = _np.asarray(get_parameter("scale", shape, init=initializers.Constant(1.0)))
        offset = _np.asarray(get_parameter("offset", shape, init=initializers.Constant(0.0)))
        mean = _np.asarray(self.mean_ema.average(shape))
        var = _np.asarray(self.var_ema.average(shape))
        inv = (scale * (_np.float32(1.0) / _np.sqrt(var + _np.float32(self.eps)))).astype(_np.float32)
        return asarr(((x - mean) * inv + offset).astype(_np.float32))


class Flatten(Module):
    def __init__(self, preserve_dims=1, name=None):
        super().__init__(name=name or "flatten")

    def __call__(self, x):
        x = _np.asarray(x)
        return asarr(x.reshape(x.shape[0], -1))


class _Transformed:
    def __init__(self, f):
        self.f = f

    def init(self, rng, *a, **k):
        fr = _Frame({}, {}, 0 if rng is None else int(_np.asarray(rng).sum()), True)
        _frames.append(fr)
        try:
            self.f(*a, **k)
        finally:
            _frames.pop()
        return fr.params, fr.state

    def apply(self, params, state, rng, *a, **k):
        fr = _Frame(params, {m: dict(v) for m, v in state.items()}, 0, False)
        _frames.append(fr)
        try:
            out = self.f(*a, **k)
        finally:
            _frames.pop()
        return out, fr.state


def transform_with_state(f):
    return _Transformed(f)


def without_apply_rng(t):
    class _NoRng:
        init = t.init

        @staticmethod
        def apply(params, state, *a, **k):
            return t.apply(params, state, None, *a, **k)

    return _NoRng()
