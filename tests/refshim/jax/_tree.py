"""Minimal pytree helpers for the stand-in (tuples, lists, dicts, None, arrays)."""
import numpy as _np


def tree_stack(items):
    """[tree_0, tree_1, ...] -> tree of arrays stacked on a new leading axis (what scan / vmap return)."""
    from . import numpy as jnp
    if not items:
        return None
    head = items[0]
    if head is None:
        return None
    if isinstance(head, (tuple, list)):
        return type(head)(tree_stack([it[k] for it in items]) for k in range(len(head)))
    if isinstance(head, dict):
        return {k: tree_stack([it[k] for it in items]) for k in head}
    return _np.stack([_np.asarray(it) for it in items], axis=0).view(jnp.Array)


def tree_index(tree, axis, i):
    from . import numpy as jnp
    if tree is None:
        return None
    if isinstance(tree, (tuple, list)):
        axes = axis if isinstance(axis, (tuple, list)) else [axis] * len(tree)
        return type(tree)(tree_index(t, a, i) for t, a in zip(tree, axes))
    if axis is None:
        return tree
    a = _np.asarray(tree)
    out = _np.take(a, i, axis=axis)
    return out.view(jnp.Array) if isinstance(out, _np.ndarray) and out.ndim else (out[()] if isinstance(out, _np.ndarray) else out)


def tree_len(tree, axis):
    if tree is None or axis is None:
        return None
    if isinstance(tree, (tuple, list)):
        axes = axis if isinstance(axis, (tuple, list)) else [axis] * len(tree)
        for t, a in zip(tree, axes):
            n = tree_len(t, a)
            if n is not None:
                return n
        return None
    return _np.asarray(tree).shape[axis]
