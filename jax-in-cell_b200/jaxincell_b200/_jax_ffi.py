"""`jax.ffi` binding of the hot path: registers the XLA FFI handler of csrc/jic_xla_ffi.cc and wraps `jax.ffi.ffi_call`.

This is the drop-in for the scan block of the reference, jaxincell/_simulation.py:216-257, usable under `jax.jit`:

    from jaxincell_b200 import _jax_ffi
    hist, fields0, v_init = _jax_ffi.boris_run(positions, velocities, ext_E, ext_B, species=[(count, q*w, m*w, q/m), ...], n_steps=T,
                                               length=..., dx=dx, dt=dt, grid=grid, bcs=(pbl, pbr, fbl, fbr), solver=solver_parameters)

It needs jax and `libjic_b200_ffi.so` (`python jax-in-cell_b200/build.py --ffi` in an environment that has jax: the XLA headers come
from `jax.ffi.include_dir()`).  Neither exists in the build image of this repository, so importing this module works but `register()`
raises; the ctypes route (`_engine.py`) is the one the tests exercise.  Autodiff through the call is out of scope (BASELINE.json)."""
from __future__ import annotations

import ctypes
import os

import numpy as np

from ._lib import JicError, LIB_PATH

FFI_LIB_PATH = os.environ.get("JIC_B200_FFI_LIB") or os.path.join(os.path.dirname(LIB_PATH), "libjic_b200_ffi.so")
TARGET = "jic_boris_run"
STEP_TARGET = "jic_boris_step"
_registered = False


def register():
    """jax.ffi.register_ffi_target(TARGET, capsule of the handler symbol, platform="CUDA") -- once per process."""
    global _registered
    if _registered:
        return
    try:
        import jax
        import jax.ffi  # noqa: F401
    except ImportError as e:
        raise JicError("the jax.ffi binding needs jax; use jaxincell_b200.simulate_host / HotPath (ctypes) without it") from e
    if not os.path.exists(FFI_LIB_PATH):
        raise JicError(f"{FFI_LIB_PATH} is missing: run `python jax-in-cell_b200/build.py --ffi` (needs jax's XLA FFI headers)")
    ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    lib = ctypes.CDLL(FFI_LIB_PATH)
    if not lib.jic_xla_ffi_available():
        raise JicError("libjic_b200_ffi.so was built without the XLA FFI headers")
    for target in (TARGET, STEP_TARGET):
        jax.ffi.register_ffi_target(target, jax.ffi.pycapsule(getattr(lib, target)), platform="CUDA")
    _registered = True


def boris_run(positions, velocities, external_E, external_B, *, species, n_steps, length, dx, dt, grid, bcs=(0, 0, 0, 0), solver=None,
              length_y=0.0, length_z=0.0, particle_history=True, engine=None):
    """One custom call for start-up + `n_steps` steps.  positions, velocities: (N,3) jax arrays at t = 0; external fields (G,3) float32.

    Returns ((positions, velocities, E, B, J, rho) histories as `lax.scan` would stack them, (E0, B0), initial_velocities);
    positions / velocities histories and initial_velocities are empty arrays when particle_history=False."""
    import jax
    import jax.numpy as jnp
    register()
    solver = dict(solver or {})
    N, G, T = positions.shape[0], external_E.shape[0], int(n_steps)
    real = positions.dtype
    hist_n = N if particle_history else 0
    if engine is None:
        engine = 0 if particle_history or N < 1_000_000 else 1
    f = lambda *shape: jax.ShapeDtypeStruct(shape, real)  # noqa: E731
    results = (f(T, G, 3), f(T, G, 3), f(T, G, 3), f(T, G), f(T, hist_n, 3), f(T, hist_n, 3), f(G, 3), f(G, 3), f(hist_n, 3))
    counts, q, m, qm = zip(*species)
    grid = np.asarray(grid, np.float64)
    call = jax.ffi.ffi_call(TARGET, results)
    E, B, J, rho, x, v, E0, B0, v_init = call(
        positions, jnp.asarray(velocities, real), jnp.asarray(external_E, jnp.float32), jnp.asarray(external_B, jnp.float32),
        n_steps=np.int64(T), length=np.float64(length), length_y=np.float64(length_y), length_z=np.float64(length_z),
        dx=np.float64(dx), dt=np.float64(dt), grid_first=np.float64(grid[0]), grid_last=np.float64(grid[-1]),
        particle_bc_left=np.int64(bcs[0]), particle_bc_right=np.int64(bcs[1]), field_bc_left=np.int64(bcs[2]), field_bc_right=np.int64(bcs[3]),
        filter_passes=np.int64(solver.get("filter_passes", 5)), filter_alpha=np.float64(solver.get("filter_alpha", 0.5)),
        filter_strides=np.asarray(solver.get("filter_strides", (1, 2, 4)), np.int64), relativistic=bool(solver.get("relativistic", False)),
        field_solver=np.int64(solver.get("field_solver", 0)), time_evolution_algorithm=np.int64(solver.get("time_evolution_algorithm", 0)),
        cn_substeps=np.int64(solver.get("number_of_particle_substeps_implicit_CN", 2)),
        cn_max_iterations=np.int64(solver.get("max_number_of_Picard_iterations_implicit_CN", 20)),
        cn_tolerance=np.float64(solver.get("tolerance_Picard_iterations_implicit_CN", 1e-6)), engine=np.int64(engine),
        species_count=np.asarray(counts, np.int64), species_charge=np.asarray(q, np.float64), species_mass=np.asarray(m, np.float64),
        species_charge_to_mass=np.asarray(qm, np.float64))
    return (x, v, E, B, J, rho), (E0, B0), v_init


def boris_step(carry, step_index, solver_parameters, external_field_parameters, dx, dt, grid, box_size, particle_BC_left, particle_BC_right,
               field_BC_left, field_BC_right, field_solver=0, *, species):
    """`Boris_step` of the reference (jaxincell/_algorithms.py:17-95) as one custom call: same positional signature, same carry and
    step_data, so it can replace `step_func` in the reference's own `lax.scan` (_simulation.py:232-253) under `jax.jit`.  The one extra,
    static argument is the species table [(count, q*w, m*w, q/m), ...] (the carry's per-particle q, m, q/m arrays are traced values and
    cannot parameterise the kernel).  q and q/m of particles the step absorbed are zeroed from the returned `alive` mask."""
    import jax
    import jax.numpy as jnp
    register()
    E, B, x_minus, x_n, x_plus, v_n, qs, ms, q_ms = carry
    N, G, real = x_plus.shape[0], E.shape[0], x_plus.dtype
    f = lambda *shape: jax.ShapeDtypeStruct(shape, real)  # noqa: E731
    results = (f(G, 3), f(G, 3), f(G, 3), f(G), f(N, 3), f(N, 3), f(N, 3), jax.ShapeDtypeStruct((N,), jnp.uint8))
    counts, q, m, qm = zip(*species)
    grid = np.asarray(grid, np.float64)
    sol = solver_parameters
    E1, B1, J, rho, x1, x_half, v1, alive = jax.ffi.ffi_call(STEP_TARGET, results)(
        E, B, x_minus, x_n, x_plus, v_n, jnp.asarray(external_field_parameters["external_electric_field"], jnp.float32),
        jnp.asarray(external_field_parameters["external_magnetic_field"], jnp.float32),
        length=np.float64(box_size[0]), length_y=np.float64(box_size[1]), length_z=np.float64(box_size[2]), dx=np.float64(dx), dt=np.float64(dt),
        grid_first=np.float64(grid[0]), grid_last=np.float64(grid[-1]), particle_bc_left=np.int64(particle_BC_left),
        particle_bc_right=np.int64(particle_BC_right), field_bc_left=np.int64(field_BC_left), field_bc_right=np.int64(field_BC_right),
        filter_passes=np.int64(sol["filter_passes"]), filter_alpha=np.float64(sol["filter_alpha"]),
        filter_strides=np.asarray(sol["filter_strides"], np.int64), relativistic=bool(sol["relativistic"]), field_solver=np.int64(field_solver),
        species_count=np.asarray(counts, np.int64), species_charge=np.asarray(q, np.float64), species_mass=np.asarray(m, np.float64),
        species_charge_to_mass=np.asarray(qm, np.float64))
    keep = (alive != 0).reshape((N,) + (1,) * (qs.ndim - 1))
    new_carry = (E1, B1, x_plus, x1, x_half, v1, jnp.where(keep, qs, 0), ms, jnp.where(keep, q_ms, 0))
    return new_carry, (x1, v1, E1, B1, J, rho)
