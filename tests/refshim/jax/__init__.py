"""NumPy stand-in for the parts of ``jax`` the reference imports (TEST INFRASTRUCTURE ONLY, see tests/refshim/README.md).

Not JAX: no tracing, no autodiff, no XLA.  ``vmap`` is a Python loop, ``jit`` the identity."""
import types as _types

from . import _tree
from . import numpy  # noqa: F401
from . import lax  # noqa: F401
from . import random  # noqa: F401

__version__ = "0.0.0+numpy-standin"


def jit(fun=None, **kw):
    if fun is None:
        return lambda f: f
    return fun


def vmap(fun, in_axes=0, out_axes=0, **kw):
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            n = _tree.tree_len(a, ax)
            if n is not None:
                break
        outs = [fun(*[_tree.tree_index(a, ax, i) for a, ax in zip(args, axes)]) for i in range(n)]
        if n == 0:
            raise ValueError("stand-in vmap over an empty axis")
        return _tree.tree_stack(outs)
    return mapped


def _unavailable(name):
    def fail(*a, **k):
        raise NotImplementedError(f"jax.{name} is not available in the NumPy stand-in")
    return fail


grad = _unavailable("grad")
value_and_grad = _unavailable("value_and_grad")
jacfwd = _unavailable("jacfwd")
jacrev = _unavailable("jacrev")


class _Config:
    def update(self, *a, **k):
        pass


config = _Config()

debug = _types.ModuleType("jax.debug")
debug.print = lambda fmt, *a, **k: print(fmt.format(*a, **{n: v for n, v in k.items() if n != 'ordered'}))
import sys as _sys  # noqa: E402
_sys.modules["jax.debug"] = debug

tree_util = _types.ModuleType("jax.tree_util")
_sys.modules["jax.tree_util"] = tree_util


def effects_barrier():
    pass


def block_until_ready(x):
    return x


def devices(*a):
    return ["cpu-numpy-standin"]
