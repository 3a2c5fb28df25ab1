"""Minimal numpy-backed stand-in for the `jax` package.

TEST INFRASTRUCTURE ONLY.  JAX is not installed in the build image (no network), and the
reference (wliverno/GauNEGF) uses JAX purely as "numpy with jit / vmap / while_loop".  This
shim lets the UNMODIFIED reference source under /root/reference import and run on numpy +
LAPACK so that (a) golden vectors can be generated from the reference itself
(tests/golden/make_golden.py) and (b) the numpy oracle in oracle/negf_oracle.py can be pinned
against it.  A sequentially executed vmap/while_loop computes, element by element, exactly what
JAX's batched versions compute (a vmapped while_loop freezes converged lanes).

Surface covered = what the reference's hot-path modules touch.
Nothing under gaunegf_b200/ imports this.
"""
import numpy as _np

from . import numpy as numpy  # noqa: F401  (jax.numpy)
from . import lax as lax      # noqa: F401
from .numpy import _wrap


class _Config:
    def update(self, *a, **k):
        return None


config = _Config()


def jit(fun=None, static_argnums=None, static_argnames=None, **_kw):
    if fun is None:
        return lambda f: f
    return fun


def vmap(fun, in_axes=0, out_axes=0):
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = _np.shape(a)[ax]
                break
        outs = []
        for i in range(n):
            call = [a if ax is None else _wrap(_np.take(_np.asarray(a), i, axis=ax))
                    for a, ax in zip(args, axes)]
            outs.append(fun(*call))
        if n and isinstance(outs[0], tuple):
            return tuple(_wrap(_np.stack([_np.asarray(o[j]) for o in outs]))
                         for j in range(len(outs[0])))
        return _wrap(_np.stack([_np.asarray(o) for o in outs]))
    return mapped


def block_until_ready(x):
    return x


def devices(*a, **k):
    return ["cpu:0 (numpy shim)"]


def clear_caches():
    return None
