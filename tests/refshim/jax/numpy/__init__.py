"""NumPy stand-in for ``jax.numpy`` (TEST INFRASTRUCTURE ONLY, see tests/refshim/README.md).

Semantics of XLA that differ from NumPy and that the reference relies on are reproduced here:
  * integer gathers ``a[i]`` with out-of-range ``i`` clamp (negative indices wrap once, Python style, then clamp);
  * ``a.at[i].set/add`` drop out-of-range updates and accumulate duplicates;
  * everything computes in float64 (the reference enables x64, jaxincell/_simulation.py:32).
"""
import sys
import types

import numpy as _np

pi = _np.pi
inf = _np.inf
nan = _np.nan
newaxis = None
float32, float64, int32, int64, bool_, complex128 = _np.float32, _np.float64, _np.int32, _np.int64, _np.bool_, _np.complex128
ndarray = _np.ndarray
integer, floating = _np.integer, _np.floating


def _is_int_index(k):
    return isinstance(k, (int, _np.integer)) or (isinstance(k, _np.ndarray) and k.dtype.kind in "iu")


def _clamp(k, n):
    k = _np.asarray(k)
    k = _np.where(k < 0, k + n, k)
    return _np.clip(k, 0, n - 1)


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, key):
        return _AtKey(self.arr, key)


class _AtKey:
    def __init__(self, arr, key):
        self.arr, self.key = arr, key

    def _valid(self):
        """Integer(-array) keys on axis 0: mask of in-range entries and the wrapped key (XLA drops out-of-range scatters)."""
        key = self.key
        if _is_int_index(key):
            n = self.arr.shape[0]
            k = _np.asarray(key)
            k = _np.where(k < 0, k + n, k)
            return (k >= 0) & (k < n), k
        return None, key

    def set(self, value):
        out = _np.array(_strip(self.arr), copy=True)
        ok, k = self._valid()
        if ok is None:
            out[k] = value
        elif k.ndim == 0:
            if ok:
                out[int(k)] = value
        else:
            value = _np.broadcast_to(_np.asarray(value), k.shape + out.shape[1:])
            out[k[ok]] = value[ok]
        return out.view(Array)

    def add(self, value):
        out = _np.array(_strip(self.arr), copy=True)
        ok, k = self._valid()
        if ok is None:
            _np.add.at(out, k, value)
        elif k.ndim == 0:
            if ok:
                out[int(k)] += value
        else:
            value = _np.broadcast_to(_np.asarray(value), k.shape + out.shape[1:])
            _np.add.at(out, k[ok], value[ok])
        return out.view(Array)


class Array(_np.ndarray):
    """ndarray with jax's ``.at`` and clamped integer gathers."""

    @property
    def at(self):
        return _At(self)

    def __getitem__(self, key):
        if self.ndim:
            if _is_int_index(key):
                key = _clamp(key, self.shape[0])
                if key.ndim == 0:
                    key = int(key)
            elif isinstance(key, tuple) and key and all(_is_int_index(k) for k in key) and len(key) <= self.ndim:
                key = tuple(_clamp(k, n) for k, n in zip(key, self.shape))
                if all(k.ndim == 0 for k in key):
                    key = tuple(int(k) for k in key)
        return super().__getitem__(key)

    def __iter__(self):
        """Exactly len(self) items (the clamped __getitem__ above never raises IndexError, which is what ends the default iteration)."""
        if self.ndim == 0:
            raise TypeError("iteration over a 0-d array")
        return (self[i] for i in range(self.shape[0]))

    # jax arrays are immutable: augmented assignment rebinds the name, item assignment is an error.
    def __iadd__(self, other):
        return self + other

    def __isub__(self, other):
        return self - other

    def __imul__(self, other):
        return self * other

    def __itruediv__(self, other):
        return self / other

    def __ifloordiv__(self, other):
        return self // other

    def __ipow__(self, other):
        return self ** other

    def __setitem__(self, key, value):
        raise TypeError("stand-in arrays are immutable like jax arrays; use .at[...].set(...)")

    def block_until_ready(self):
        return self

    def __hash__(self):  # jax arrays are unhashable too, but 0-d ones appear as dict values only
        raise TypeError("unhashable")


def _wrap(x):
    if isinstance(x, _np.ndarray):
        return x.view(Array)
    if isinstance(x, (tuple, list)):
        return type(x)(_wrap(v) for v in x)
    return x


def _strip(x):
    """Array -> plain ndarray view, so that NumPy's own helpers may write into their temporaries."""
    if isinstance(x, Array):
        return x.view(_np.ndarray)
    if isinstance(x, (tuple, list)):
        return type(x)(_strip(v) for v in x)
    return x


def _wrapped(fn):
    def call(*a, **k):
        return _wrap(fn(*_strip(a), **{n: _strip(v) for n, v in k.items()}))
    call.__name__ = getattr(fn, "__name__", "fn")
    return call


def array(x, dtype=None, **kw):
    return _np.array(_strip(x), dtype=dtype).view(Array)


def asarray(x, dtype=None, **kw):
    return _np.array(_strip(x), dtype=dtype).view(Array)


def roll(a, shift, axis=None):
    if isinstance(shift, _np.ndarray):
        shift = int(shift)
    return _np.roll(_strip(a), shift, axis=axis).view(Array)


def where(cond, x=None, y=None):
    if x is None:
        return _wrap(_np.where(cond))
    return _np.asarray(_np.where(_strip(cond), _strip(x), _strip(y))).view(Array)


def select(condlist, choicelist, default=0):
    out = _np.asarray(default, dtype=float)
    for c, v in reversed(list(zip(condlist, choicelist))):
        out = _np.where(_strip(c), _strip(v), out)
    return _np.asarray(out).view(Array)


def arange(*a, **k):
    return _np.arange(*[int(v) if isinstance(v, _np.ndarray) and v.ndim == 0 and v.dtype.kind in "iu" else v for v in a], **k).view(Array)


def linspace(start, stop, num=50, endpoint=True, **kw):
    """jnp.linspace evaluates start*(1-t) + stop*t with t = i/(num-1) and pins the last point (jax/_src/numpy/lax_numpy.py)."""
    num = int(num)
    if not endpoint or num < 2:
        return _np.linspace(start, stop, num, endpoint=endpoint).view(Array)
    t = _np.arange(num - 1, dtype=float) / (num - 1)
    start, stop = _np.asarray(start, float), _np.asarray(stop, float)
    body = start * (1.0 - t) + stop * t
    return _np.concatenate([body, [stop]]).view(Array)


def dot(a, b):
    return _wrap(_np.dot(_strip(a), _strip(b)))


def cross(a, b, **k):
    return _np.cross(_strip(a), _strip(b), **k).view(Array)


_fft = types.ModuleType("jax.numpy.fft")
for _n in ("fft", "ifft", "fftfreq", "rfft", "irfft", "fftshift"):
    setattr(_fft, _n, _wrapped(getattr(_np.fft, _n)))
fft = _fft
sys.modules["jax.numpy.fft"] = _fft

_linalg = types.ModuleType("jax.numpy.linalg")
for _n in ("solve", "norm", "inv", "det"):
    setattr(_linalg, _n, _wrapped(getattr(_np.linalg, _n)))
linalg = _linalg
sys.modules["jax.numpy.linalg"] = _linalg


def __getattr__(name):
    obj = getattr(_np, name)
    if callable(obj) and not isinstance(obj, type):
        return _wrapped(obj)
    return obj
