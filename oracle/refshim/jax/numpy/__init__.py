"""jax.numpy stand-in: the numpy namespace, with array constructors returning an ndarray
subclass that carries the functional-update accessor `.at[idx].set/.add` (test infrastructure)."""
import numpy as _np
from numpy import *  # noqa: F401,F403
from numpy import testing  # noqa: F401
from . import linalg  # noqa: F401


class _AtIndexer:
    def __init__(self, arr, idx):
        self._arr, self._idx = arr, idx

    def set(self, v):
        out = _np.array(self._arr, copy=True).view(JArr)
        out[self._idx] = v
        return out

    def add(self, v):
        out = _np.array(self._arr, copy=True).view(JArr)
        out[self._idx] += v
        return out


class _At:
    def __init__(self, arr):
        self._arr = arr

    def __getitem__(self, idx):
        return _AtIndexer(self._arr, idx)


class JArr(_np.ndarray):
    @property
    def at(self):
        return _At(self)


def _wrap(x):
    if isinstance(x, _np.ndarray) and not isinstance(x, JArr):
        return x.view(JArr)
    return x


def _wrapped(fn):
    def f(*a, **k):
        return _wrap(fn(*a, **k))
    f.__name__ = getattr(fn, "__name__", "f")
    return f


for _name in ("array", "asarray", "zeros", "zeros_like", "ones", "ones_like", "eye", "diag",
              "kron", "concatenate", "arange", "stack", "where", "sum", "conj", "real", "imag",
              "abs", "maximum", "power", "dot", "matmul", "linspace", "copy", "mean"):
    globals()[_name] = _wrapped(getattr(_np, _name))

complex128 = _np.complex128
float64 = _np.float64
inf = _np.inf
pi = _np.pi
ndarray = _np.ndarray
