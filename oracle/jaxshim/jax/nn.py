import numpy as _np

from ._core import asarr


def relu(x):
    return asarr(_np.maximum(_np.asarray(x), 0))


def softmax(x, axis=-1):
    x = _np.asarray(x)
    e = _np.exp(x - x.max(axis=axis, keepdims=True))
    return asarr(e / e.sum(axis=axis, keepdims=True))


def softplus(x):
    return asarr(_np.logaddexp(_np.asarray(x), 0))
