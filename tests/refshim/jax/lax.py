"""Python control flow standing in for ``jax.lax`` (TEST INFRASTRUCTURE ONLY, see tests/refshim/README.md)."""
import numpy as _np

from . import numpy as jnp
from ._tree import tree_stack


def cond(pred, true_fun, false_fun, *operands, operand=None, **kw):
    if not operands:
        operands = (operand,)
    return true_fun(*operands) if bool(pred) else false_fun(*operands)


def scan(f, init, xs=None, length=None, **kw):
    carry = init
    n = length if xs is None else len(xs)
    ys = []
    for i in range(int(n)):
        x = None if xs is None else xs[i]
        carry, y = f(carry, x)
        ys.append(y)
    return carry, tree_stack(ys)


def while_loop(cond_fun, body_fun, init_val):
    val = init_val
    while bool(cond_fun(val)):
        val = body_fun(val)
    return val


def fori_loop(lower, upper, body_fun, init_val):
    val = init_val
    for i in range(int(lower), int(upper)):
        val = body_fun(i, val)
    return val


def slice(operand, start_indices, limit_indices, strides=None):  # noqa: A001
    key = tuple(_np.s_[int(a):int(b)] for a, b in zip(start_indices, limit_indices))
    return _np.asarray(operand)[key].view(jnp.Array)


def dynamic_update_slice(operand, update, start_indices):
    """XLA clamps the start so that the update fits."""
    out = _np.array(_np.asarray(operand).view(_np.ndarray), copy=True)
    update = _np.asarray(update)
    key = []
    for s, n, m in zip(start_indices, out.shape, update.shape):
        s = min(max(int(s), 0), n - m)
        key.append(_np.s_[s:s + m])
    out[tuple(key)] = update
    return out.view(jnp.Array)


def stop_gradient(x):
    return x
