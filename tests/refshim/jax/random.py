"""``jax.random`` stand-in on top of oracle/sampling.py (Threefry-2x32 restatement; TEST INFRASTRUCTURE ONLY).

Only integer seeds through ``PRNGKey(seed)`` and 1-D float64 ``uniform`` / ``normal`` draws, which is all the reference uses
(jaxincell/_state_initialization.py:57-76)."""
import os
import sys

import numpy as _np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
from oracle import sampling as _S  # noqa: E402

from . import numpy as jnp  # noqa: E402


class _Key:
    def __init__(self, seed):
        self.seed = int(seed)


def PRNGKey(seed):
    return _Key(seed)


key = PRNGKey


def uniform(key, shape=(), dtype=float, minval=0.0, maxval=1.0):
    n = int(_np.prod(shape)) if shape else 1
    out = _S.uniform64(key.seed, n, float(minval), float(maxval))
    return _np.asarray(out).reshape(shape).view(jnp.Array)


def normal(key, shape=(), dtype=float):
    n = int(_np.prod(shape)) if shape else 1
    return _np.asarray(_S.normal64(key.seed, n)).reshape(shape).view(jnp.Array)
