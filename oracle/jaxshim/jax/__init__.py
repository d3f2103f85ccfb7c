"""numpy stand-in for the slice of `jax` used by the reference hot-path modules (see ../README.md)."""
import numpy as _np
from . import numpy, lax, random, nn, tree  # noqa: F401
from ._core import Arr as Array, tree_map, tree_leaves, vmap, asarr  # noqa: F401

Device = object


def jit(f, *a, **k):
    return f


def local_devices():
    return [object()]


def pmap(f, *a, **k):
    return f
