"""Keys are numpy Generators' seeds; none of the pinned paths consumes randomness
 except Subleq._init's task choice, which make_golden.py fixes to a single task."""
import numpy as _np

from ._core import asarr


def PRNGKey(seed):
    return asarr(_np.array([0, seed], dtype=_np.uint32))


key = PRNGKey


def split(key, num=2):
    k = _np.asarray(key).astype(_np.uint64)
    base = int(k[0]) * 1000003 + int(k[1]) * 7919 + 1
    return asarr(_np.array([[(base + i) >> 32 & 0xFFFFFFFF, (base + i * 2654435761) & 0xFFFFFFFF] for i in range(num)], dtype=_np.uint32))


def _rng(key):
    k = _np.asarray(key).astype(_np.uint64)
    return _np.random.default_rng(int(k[0]) * 4294967296 + int(k[1]))


def choice(key, a):
    a = _np.asarray(a)
    return asarr(a[_rng(key).integers(0, a.shape[0])])


def bernoulli(key, p=0.5, shape=None):
    return asarr(_rng(key).random(shape) < p)


def gumbel(key, shape, dtype=_np.float32):
    return asarr(_rng(key).gumbel(size=shape).astype(_np.float32))
