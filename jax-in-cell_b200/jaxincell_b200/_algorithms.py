"""`Boris_step` and `CN_step` with the reference's signatures and carries (jaxincell/_algorithms.py:17-95, :100-241), executed by the library:

    carry, step_data = Boris_step(carry, step_index, solver_parameters, external_field_parameters, dx, dt, grid, box_size,
                                  particle_BC_left, particle_BC_right, field_BC_left, field_BC_right, field_solver)

    carry     = (E, B, x_{n-1/2}, x_n, x_{n+1/2}, v_n, q, m, q/m)          (_simulation.py:228-231; q, m, q/m are (N,1) or (N,))
    step_data = (x_{n+1}, v_{n+1}, E^{n+1}, B^{n+1}, J, rho)                (_algorithms.py:93)

This is the step-granularity drop-in: the name the reference's tests monkeypatch at (`_algorithms.Boris_step`) and the body of its
`lax.scan`.  Every call loads the carry into a cached device context (jic_load_carry), advances one step (jic_run) and reads the new
carry back, so it costs host<->device traffic of the whole particle state per step -- use `Simulation.run()` / `HotPath.run(T)` (the
whole scan in one call) for anything but step-level interoperability and testing.  NumPy in, NumPy out; no autodiff.

Species: the library works on contiguous blocks of identical particles, which is how the reference lays them out
(_state_initialization.py:242-261).  The blocks are recovered from the carry's mass array (masses are never modified, charges and
q/m of absorbed particles are zeroed: _boundary_conditions.py:96-102); a block's charge and q/m come from any particle of the block
that is still alive."""
from __future__ import annotations

import numpy as np

from ._lib import JIC_MAX_SPECIES, JicError

_contexts = {}


def species_blocks(qs, ms, q_ms):
    """[(count, q, m, q/m)] of the contiguous runs of equal mass in the carry.  Absorbed particles (q = q/m = 0) inherit the values of
    their block; a block with no survivor keeps q = q/m = 0 (it never deposits or moves again)."""
    qs, ms, q_ms = (np.asarray(a, np.float64).reshape(-1) for a in (qs, ms, q_ms))
    if not (len(qs) == len(ms) == len(q_ms)):
        raise JicError("q, m, q/m of the carry must have the same length")
    if len(ms) == 0:
        return [dict(count=0, q=0.0, m=1.0, qm=0.0)]
    cuts = np.flatnonzero(ms[1:] != ms[:-1]) + 1
    starts = np.concatenate([[0], cuts])
    ends = np.concatenate([cuts, [len(ms)]])
    if len(starts) > JIC_MAX_SPECIES:
        raise JicError(f"the carry has {len(starts)} runs of distinct mass; the library takes at most {JIC_MAX_SPECIES} species blocks")
    blocks = []
    for a, b in zip(starts, ends):
        alive = np.flatnonzero(qs[a:b] != 0.0)
        if len(alive):
            q, qm = float(qs[a + alive[0]]), float(q_ms[a + alive[0]])
            if np.any(qs[a:b][alive] != q) or np.any(q_ms[a:b][alive] != qm):
                raise JicError("particles of equal mass with different charge inside one block: not a species block of the reference")
        else:
            q, qm = 0.0, 0.0
        blocks.append(dict(count=int(b - a), q=q, m=float(ms[a]), qm=qm))
    return blocks


def _context(key, make):
    hp = _contexts.get(key)
    if hp is None:
        if len(_contexts) >= 4:  # a handful of live configurations at most: device memory is not a cache to grow
            _contexts.pop(next(iter(_contexts))).close()
        hp = _contexts[key] = make()
    return hp


def Boris_step(carry, step_index, solver_parameters, external_field_parameters, dx, dt, grid, box_size,
               particle_BC_left, particle_BC_right, field_BC_left, field_BC_right, field_solver=0):
    from ._engine import HotPath
    E, B, x_minus, x_n, x_plus, v_n, qs, ms, q_ms = carry
    grid = np.asarray(grid, np.float64)
    G = len(grid)
    blocks = species_blocks(qs, ms, q_ms)
    sol = solver_parameters
    ext_E = np.asarray(external_field_parameters["external_electric_field"], np.float32)
    ext_B = np.asarray(external_field_parameters["external_magnetic_field"], np.float32)
    length = float(box_size[0])
    key = (G, float(dx), float(dt), tuple(float(b) for b in box_size), int(particle_BC_left), int(particle_BC_right), int(field_BC_left),
           int(field_BC_right), int(sol["filter_passes"]), float(sol["filter_alpha"]), tuple(int(s) for s in sol["filter_strides"]),
           bool(sol["relativistic"]), int(field_solver), tuple((b["count"], b["q"], b["m"], b["qm"]) for b in blocks))
    hp = _context(key, lambda: HotPath(species=blocks, length=length, length_y=float(box_size[1]), length_z=float(box_size[2]), G=G, dt=float(dt),
                                       pbl=int(particle_BC_left), pbr=int(particle_BC_right), fbl=int(field_BC_left), fbr=int(field_BC_right),
                                       filter_passes=int(sol["filter_passes"]), filter_alpha=float(sol["filter_alpha"]),
                                       filter_strides=tuple(int(s) for s in sol["filter_strides"]), relativistic=bool(sol["relativistic"]),
                                       engine="indexed", track_yz=True, field_solver=int(field_solver)))
    if abs(hp.params.dx - float(dx)) > 1e-14 * abs(float(dx)):
        raise JicError("dx is not box_size[0] / len(grid)")
    # The library marks an absorbed particle by its parked position (outside the box), the reference by q == 0
    # (_boundary_conditions.py:40-56 does both at once).  The two agree for every carry the reference's own steps produce as long as
    # a reflected particle cannot overshoot the far wall in one step (CFL < number of cells); a hand-made carry that breaks the
    # equivalence is refused rather than silently revived or killed.
    q_flat = np.asarray(qs, np.float64).reshape(-1)
    xp = np.asarray(x_plus, np.float64).reshape(-1, 3)[:, 0]
    inside = (xp >= -length / 2) & (xp <= length / 2)
    if np.any((q_flat == 0.0) & inside):
        raise JicError("carry holds particles with q == 0 inside the box: the library identifies absorbed particles by their parked position")
    if np.any((q_flat != 0.0) & ~inside):
        raise JicError("carry holds charged particles outside the box (a reflection that overshot the far wall: CFL >= number of cells?)")
    hp.set_external_fields(ext_E, ext_B)
    hp.load_carry(E, B, x_minus, x_n, x_plus, v_n)
    out = hp.run(1, particles=True)
    E1, B1, J1, rho1 = (out[k][0].cpu().numpy() for k in ("electric_field", "magnetic_field", "current_density", "charge_density"))
    x1, v1 = out["positions"][0].cpu().numpy(), out["velocities"][0].cpu().numpy()
    x_half, _, alive = (t.cpu().numpy() for t in hp.particles())
    alive = alive.astype(bool)
    shape = np.asarray(qs).shape
    qs1 = np.where(alive, np.asarray(qs, np.float64).reshape(-1), 0.0).reshape(shape)
    qms1 = np.where(alive, np.asarray(q_ms, np.float64).reshape(-1), 0.0).reshape(shape)
    new_carry = (E1, B1, np.asarray(x_plus), x1, x_half, v1, qs1, ms, qms1)
    return new_carry, (x1, v1, E1, B1, J1, rho1)


def CN_step(carry, step_index, solver_parameters, dx, dt, grid, box_size, particle_BC_left, particle_BC_right, field_BC_left, field_BC_right,
            num_substeps):
    """`CN_step` with the reference's signature and carry (jaxincell/_algorithms.py:100-241): carry = (E, B, x_n, v_n, q, m, q/m),
    step_data = (x_{n+1}, v_{n+1}, E^{n+1}, B^{n+1}, J, rho).  The charges of the carry pass through unchanged, as in the reference."""
    from ._engine import HotPath
    E, B, x_n, v_n, qs, ms, q_ms = carry
    grid = np.asarray(grid, np.float64)
    G = len(grid)
    blocks = species_blocks(qs, ms, q_ms)
    sol = solver_parameters
    key = ("cn", G, float(dx), float(dt), tuple(float(b) for b in box_size), int(particle_BC_left), int(particle_BC_right), int(field_BC_left),
           int(field_BC_right), int(num_substeps), int(sol["max_number_of_Picard_iterations_implicit_CN"]),
           float(sol["tolerance_Picard_iterations_implicit_CN"]), tuple((b["count"], b["q"], b["m"], b["qm"]) for b in blocks))
    hp = _context(key, lambda: HotPath(species=blocks, length=float(box_size[0]), length_y=float(box_size[1]), length_z=float(box_size[2]), G=G,
                                       dt=float(dt), pbl=int(particle_BC_left), pbr=int(particle_BC_right), fbl=int(field_BC_left),
                                       fbr=int(field_BC_right), engine="indexed", track_yz=True, time_evolution_algorithm=1,
                                       cn_substeps=int(num_substeps), cn_max_iterations=int(sol["max_number_of_Picard_iterations_implicit_CN"]),
                                       cn_tolerance=float(sol["tolerance_Picard_iterations_implicit_CN"])))
    hp.set_external_fields(None, None)
    hp.load_carry_cn(E, B, x_n, v_n, alive=np.asarray(qs, np.float64).reshape(-1) != 0.0)
    out = hp.run(1, particles=True)
    E1, B1, J1, rho1 = (out[k][0].cpu().numpy() for k in ("electric_field", "magnetic_field", "current_density", "charge_density"))
    x1, v1 = out["positions"][0].cpu().numpy(), out["velocities"][0].cpu().numpy()
    return (E1, B1, x1, v1, qs, ms, q_ms), (x1, v1, E1, B1, J1, rho1)


def release_contexts():
    """Free the cached device contexts of Boris_step."""
    while _contexts:
        _contexts.popitem()[1].close()
