"""jax.lax stand-in: sequential while_loop / scan / cond (test infrastructure)."""
import numpy as _np


def while_loop(cond_fun, body_fun, init_val):
    val = init_val
    while bool(cond_fun(val)):
        val = body_fun(val)
    return val


def scan(f, init, xs, length=None):
    from .numpy import _wrap
    carry = init
    n = len(xs[0]) if isinstance(xs, (tuple, list)) else len(xs)
    ys = []
    for i in range(n):
        x = tuple(a[i] for a in xs) if isinstance(xs, (tuple, list)) else xs[i]
        carry, y = f(carry, x)
        ys.append(_np.asarray(y))
    stacked = _wrap(_np.stack(ys)) if ys else _wrap(_np.zeros((0,)))
    return carry, stacked


def cond(pred, true_fun, false_fun, *operands):
    return true_fun(*operands) if bool(pred) else false_fun(*operands)
