"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/c/jic_oracle.c, the compiled (gcc, OpenMP) restatement of the explicit Boris
step.  Same contract as ``oracle.closed_form.run``.  Only ``tests/``, ``__graft_entry__`` and ``bench.py``'s CPU legs may import this.

``build()`` compiles oracle/c/jic_oracle.c into oracle/_build/libjic_oracle.so (git-ignored; travels to the GPU box with the
snapshot, and is rebuilt there on demand since gcc is part of the image).  Parity status: see the header of the C file."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import closed_form as CF

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c", "jic_oracle.c")
LIB = os.path.join(HERE, "_build", "libjic_oracle.so")
_lib = None


class Params(C.Structure):
    _fields_ = [("G", C.c_int32), ("pbl", C.c_int32), ("pbr", C.c_int32), ("fbl", C.c_int32), ("fbr", C.c_int32),
                ("filter_passes", C.c_int32), ("n_strides", C.c_int32), ("strides", C.c_int32 * 8),
                ("relativistic", C.c_int32), ("field_solver", C.c_int32), ("n_threads", C.c_int32),
                ("L", C.c_double), ("Ly", C.c_double), ("Lz", C.c_double), ("dx", C.c_double), ("dt", C.c_double), ("filter_alpha", C.c_double)]


def build(force=False):
    """gcc -O2 -fopenmp -ffp-contract=off (no FMA contraction: the arithmetic of the NumPy oracles)."""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    tmp = LIB + f".{os.getpid()}.tmp"
    cmd = ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-Wextra", SRC, "-o", tmp, "-lm"]
    try:
        res = subprocess.run(cmd, capture_output=True, text=True)
        failure = None if res.returncode == 0 else res.stdout + res.stderr
    except OSError as e:  # no compiler on this host
        failure = str(e)
    if failure is not None:
        if os.path.exists(LIB) and not force:
            return LIB  # a prebuilt library travelled with the snapshot (copies do not keep mtimes): use it
        raise RuntimeError("gcc failed:\n" + failure)
    os.replace(tmp, LIB)
    return LIB


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        assert lib.jo_abi_version() == 2
        lib.jo_run.restype = C.c_int
        lib.jo_run.argtypes = [C.POINTER(Params), C.c_int64] + [C.c_void_p] * 9 + [C.c_int64] + [C.c_void_p] * 12
        _lib = lib
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def run(x0, v0, qs, ms, q_ms, *, length, G, dt, total_steps, box_yz=None, pbl=0, pbr=0, fbl=0, fbr=0, solver=None, ext_E=None, ext_B=None,
        keep_particles=True, threads=None):
    """Same inputs / outputs as ``closed_form.run`` (histories of x_{n+1}, v_{n+1}, E, B, J, rho; `fields`, `initial_velocities`),
    plus `x_half`, `v_final` (the carried state after the last step) and `step_seconds` (wall-clock of every step).  threads: OpenMP threads (default: all cores)."""
    lib = load()
    solver = {"filter_passes": 5, "filter_alpha": 0.5, "filter_strides": (1, 2, 4), "relativistic": False, "field_solver": 0, **(solver or {})}
    dom = CF.Domain(length, G, dt, *(box_yz if box_yz is not None else (None, None)))
    f64 = lambda a, shape: np.ascontiguousarray(np.asarray(a, np.float64).reshape(shape))  # noqa: E731
    N = len(np.asarray(x0))
    x0, v0 = f64(x0, (N, 3)), f64(v0, (N, 3))
    qs, ms, q_ms = f64(qs, (N,)), f64(ms, (N,)), f64(q_ms, (N,))
    p = Params()
    p.G, p.pbl, p.pbr, p.fbl, p.fbr = G, pbl, pbr, fbl, fbr
    strides = tuple(int(s) for s in solver["filter_strides"])
    p.filter_passes, p.n_strides, p.filter_alpha = int(solver["filter_passes"]), len(strides), float(solver["filter_alpha"])
    for i, s in enumerate(strides):
        p.strides[i] = s
    p.relativistic, p.field_solver = int(bool(solver["relativistic"])), int(solver["field_solver"])
    p.n_threads = int(threads or len(os.sched_getaffinity(0)) or 1)
    p.L, p.Ly, p.Lz, p.dx, p.dt = dom.L, dom.Ly, dom.Lz, dom.dx, dom.dt
    grid = np.ascontiguousarray(dom.grid)
    eE = None if ext_E is None else f64(np.asarray(ext_E, np.float32), (G, 3))
    eB = None if ext_B is None else f64(np.asarray(ext_B, np.float32), (G, 3))
    h = np.ascontiguousarray(CF.gauss_kernel(G, dom.dx)) if p.field_solver in (1, 3) else None
    T = int(total_steps)
    out = dict(electric_field=np.empty((T, G, 3)), magnetic_field=np.empty((T, G, 3)), current_density=np.empty((T, G, 3)),
               charge_density=np.empty((T, G)))
    if keep_particles:
        out.update(positions=np.empty((T, N, 3)), velocities=np.empty((T, N, 3)))
    E0, B0, vi, xh, vf = np.empty((G, 3)), np.empty((G, 3)), np.empty((N, 3)), np.empty((N, 3)), np.empty((N, 3))
    secs = np.zeros(T)
    rc = lib.jo_run(C.byref(p), N, _ptr(x0), _ptr(v0), _ptr(qs), _ptr(ms), _ptr(q_ms), _ptr(grid), _ptr(eE), _ptr(eB), _ptr(h), T,
                    _ptr(out["electric_field"]), _ptr(out["magnetic_field"]), _ptr(out["current_density"]), _ptr(out["charge_density"]),
                    _ptr(out.get("positions")), _ptr(out.get("velocities")), _ptr(E0), _ptr(B0), _ptr(vi), _ptr(xh), _ptr(vf), _ptr(secs))
    if rc != 0:
        raise MemoryError("jo_run: allocation failed")
    out.update(grid=dom.grid, dx=dom.dx, dt=dom.dt, initial_velocities=vi, fields=(E0, B0), x_half=xh, v_final=vf, step_seconds=secs, threads=p.n_threads)
    return out
