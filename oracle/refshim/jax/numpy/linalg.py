"""jax.numpy.linalg stand-in → numpy.linalg (LAPACK zgesv/zgetri/zheev), test infrastructure."""
import numpy as _np
from numpy.linalg import *  # noqa: F401,F403


def _w(x):
    from . import _wrap
    return _wrap(x)


def solve(a, b):
    return _w(_np.linalg.solve(_np.asarray(a), _np.asarray(b)))


def inv(a):
    return _w(_np.linalg.inv(_np.asarray(a)))


def eigh(a):
    w, v = _np.linalg.eigh(_np.asarray(a))
    return _w(w), _w(v)


def eig(a):
    w, v = _np.linalg.eig(_np.asarray(a))
    return _w(w), _w(v)


def svd(a, *args, **kw):
    r = _np.linalg.svd(_np.asarray(a), *args, **kw)
    return tuple(_w(x) for x in r) if isinstance(r, tuple) else _w(r)


def norm(a, *args, **kw):
    return _np.linalg.norm(_np.asarray(a), *args, **kw)
