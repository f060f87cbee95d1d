"""TEST INFRASTRUCTURE ONLY -- literal NumPy restatement of JAX-in-Cell's explicit (Boris) PIC step.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline/reference legs may import this.
The product path (``jax-in-cell_b200``) never does.

This module restates, expression by expression, the *semantics* of the reference hot path in fp64 NumPy,
including its O(N*G) "every particle touches every node" formulation, so that it can be pinned against the
known-answer vectors of the reference's own unit tests (tests/test_oracle_kat.py) and then used to validate
the O(N) closed form in ``oracle/closed_form.py`` and the CUDA kernels.

Parity status: PINNED -- callees by the hand-computed vectors of the reference's own unit tests (tests/test_oracle_kat.py,
SURVEY.md section 8c); the composed step / start-up / scan by running the reference's own source files on a NumPy stand-in for
jax (tests/refshim, itself pinned by 183 of the reference's 188 unit tests): tests/golden/refsrc_*.npz, reproduced here to
<= 2e-13 relative (tests/test_golden.py, tests/golden/REFERENCE_SOURCE_RUN.md).  Not pinned: XLA's floating-point evaluation
order (JAX is not installable in the build image), a round-off effect.

All ``file:line`` citations are relative to the reference checkout (``jaxincell/...``).
"""
from __future__ import annotations

import numpy as np

# jaxincell/_constants.py:1-7 -- the rounded values are part of the results.
epsilon_0 = 8.85418782e-12
mu_0 = 1.25663706e-6
speed_of_light = 2.99792458e8
elementary_charge = 1.60217663e-19
mass_electron = 9.10938371e-31
mass_proton = 1.67262193e-27
boltzmann_constant = 1.380649e-23

PERIODIC, REFLECTIVE, ABSORBING = 0, 1, 2


# ----------------------------------------------------------------------------------------------------
# particle boundaries -- jaxincell/_boundary_conditions.py:7-145
# ----------------------------------------------------------------------------------------------------
def _floor_mod(a, b):
    """XLA/NumPy float ``%`` (sign of the divisor)."""
    return np.mod(a, b)


def _bc_x_position(x, dx, grid, Lx, bc_left, bc_right):
    """x-component of the particle BC (shared by the full and the positions-only variants).

    _boundary_conditions.py:32-56 and :118-128.  Strict inequalities: x == +-L/2 is left alone.
    """
    x = np.asarray(x, dtype=np.float64)
    wrapped = _floor_mod(x + Lx / 2, Lx) - Lx / 2
    left_val = {PERIODIC: wrapped, REFLECTIVE: -Lx - x, ABSORBING: np.full_like(x, grid[0] - 1.5 * dx)}[bc_left]
    right_val = {PERIODIC: wrapped, REFLECTIVE: Lx - x, ABSORBING: np.full_like(x, grid[-1] + 3 * dx)}[bc_right]
    return np.where(x < -Lx / 2, left_val, np.where(x > Lx / 2, right_val, x))


def set_BC_positions(xs, dx, grid, Lx, Ly, Lz, bc_left, bc_right):
    """Positions-only BC, (N,3) -> (N,3).  _boundary_conditions.py:104-145."""
    xs = np.asarray(xs, dtype=np.float64)
    out = np.empty_like(xs)
    out[:, 0] = _bc_x_position(xs[:, 0], dx, grid, Lx, bc_left, bc_right)
    out[:, 1] = _floor_mod(xs[:, 1] + Ly / 2, Ly) - Ly / 2
    out[:, 2] = _floor_mod(xs[:, 2] + Lz / 2, Lz) - Lz / 2
    return out


def set_BC_particles(xs, vs, qs, ms, q_ms, dx, grid, Lx, Ly, Lz, bc_left, bc_right):
    """Full particle BC.  _boundary_conditions.py:7-102.

    qs, ms, q_ms are (N,) here (the reference carries (N,1)); masses are never modified (:102).
    """
    xs = np.asarray(xs, dtype=np.float64)
    vs = np.asarray(vs, dtype=np.float64)
    x = xs[:, 0]
    out_left = x < -Lx / 2
    out_right = x > Lx / 2
    new_xs = set_BC_positions(xs, dx, grid, Lx, Ly, Lz, bc_left, bc_right)

    flip = np.array([-1.0, 1.0, 1.0])

    def v_for(bc):
        if bc == PERIODIC:
            return vs
        if bc == REFLECTIVE:
            return vs * flip
        return np.zeros_like(vs)

    new_vs = np.where(out_left[:, None], v_for(bc_left), np.where(out_right[:, None], v_for(bc_right), vs))
    absorbed = (out_left & (bc_left == ABSORBING)) | (out_right & (bc_right == ABSORBING))
    new_qs = np.where(absorbed, 0.0, qs)
    new_qms = np.where(absorbed, 0.0, q_ms)
    return new_xs, new_vs, new_qs, ms, new_qms


# ----------------------------------------------------------------------------------------------------
# field ghost cells -- jaxincell/_boundary_conditions.py:148-247
# ----------------------------------------------------------------------------------------------------
def field_ghost_cells_E(bc_left, bc_right, E, B):
    """Ghost rows for curl E.  _boundary_conditions.py:170-178."""
    c = speed_of_light
    zero = np.zeros(3)
    L = {0: E[-1], 1: E[0], 2: np.array([0.0, -2 * c * B[0, 2] - E[0, 1], 2 * c * B[0, 1] - E[0, 2]])}.get(bc_left, zero)
    R = {0: E[0], 1: E[-1], 2: np.array([0.0, 3 * E[-1, 1] - 2 * c * B[-1, 2], 3 * E[-1, 2] + 2 * c * B[-1, 1]])}.get(bc_right, zero)
    return np.asarray(L, dtype=np.float64), np.asarray(R, dtype=np.float64)


def field_ghost_cells_B(bc_left, bc_right, B, E):
    """Ghost rows for curl B.  _boundary_conditions.py:199-207."""
    c = speed_of_light
    zero = np.zeros(3)
    L = {0: B[-1], 1: B[0], 2: np.array([0.0, 3 * B[0, 1] - (2 / c) * E[0, 2], 3 * B[0, 2] + (2 / c) * E[0, 1]])}.get(bc_left, zero)
    R = {0: B[0], 1: B[-1], 2: np.array([0.0, -(2 / c) * E[-1, 2] - B[-1, 1], (2 / c) * E[-1, 1] - B[-1, 2]])}.get(bc_right, zero)
    return np.asarray(L, dtype=np.float64), np.asarray(R, dtype=np.float64)


def field_2_ghost_cells(bc_left, bc_right, field):
    """Two left + one right ghost rows used by the gather.  _boundary_conditions.py:233-247."""
    zero = np.zeros(field.shape[1:])
    L2 = {0: field[-2], 1: field[1]}.get(bc_left, zero)
    L1 = {0: field[-1], 1: field[0]}.get(bc_left, zero)
    R = {0: field[0], 1: field[-1]}.get(bc_right, zero)
    return np.asarray(L2, dtype=np.float64), np.asarray(L1, dtype=np.float64), np.asarray(R, dtype=np.float64)


# ----------------------------------------------------------------------------------------------------
# gather + pushers -- jaxincell/_particles.py
# ----------------------------------------------------------------------------------------------------
def fields_to_particles_grid(x_n, field, dx, grid, grid_start, bc_left, bc_right):
    """Quadratic-spline gather of a (G,3) field at ONE particle position x_n = (x,y,z).

    _particles.py:29-45.  Out-of-range indices clamp, as XLA gathers do.
    """
    field = np.asarray(field, dtype=np.float64)
    grid = np.asarray(grid, dtype=np.float64)
    L2, L1, R = field_2_ghost_cells(bc_left, bc_right, field)
    padded = np.concatenate([L2[None], L1[None], field, R[None]], axis=0)  # :30-32
    x = float(np.asarray(x_n)[0])
    nodes = np.concatenate([[grid[0] - dx], grid])  # :37
    i = int(np.floor_divide(x - grid_start + dx, dx))  # :40

    def clampi(k, n):
        if k < 0:  # negative indices wrap once in XLA/NumPy-style indexing
            k += n
        return min(max(k, 0), n - 1)

    g = nodes[clampi(i, len(nodes))]
    f0 = padded[clampi(i, len(padded))]
    f1 = padded[clampi(i + 1, len(padded))]
    f2 = padded[clampi(i + 2, len(padded))]
    # :43
    return 0.5 * f0 * (0.5 + (g - x) / dx) ** 2 + f1 * (0.75 - (g - x) ** 2 / dx ** 2) + 0.5 * f2 * (0.5 - (g - x) / dx) ** 2


def rotation(dt, B, vsub, q_m):
    """Boris rotation of one velocity (3,) -- or of (N,3) velocities at once with q_m (N,1).  _particles.py:85-93."""
    B = np.asarray(B, dtype=np.float64)
    vsub = np.asarray(vsub, dtype=np.float64)
    Rvec = vsub + 0.5 * dt * q_m * np.cross(vsub, B)
    Bvec = 0.5 * q_m * dt * B
    dot = lambda a, b: np.sum(a * b, axis=-1, keepdims=a.ndim > 1)
    return (np.cross(Rvec, Bvec) + dot(Rvec, Bvec) * Bvec + Rvec) / (1 + dot(Bvec, Bvec))


def boris_step(dt, xs_half, vs, q_ms, E_at_x, B_at_x):
    """Non-relativistic Boris push for (N,3) arrays; q_ms is (N,1).  _particles.py:116-127."""
    q_ms = np.asarray(q_ms, dtype=np.float64).reshape(-1, 1)
    v_minus = vs + q_ms * E_at_x * dt / 2
    v_rot = rotation(dt, np.asarray(B_at_x, dtype=np.float64), v_minus, q_ms) if len(vs) else v_minus
    v_new = v_rot + q_ms * E_at_x * dt / 2
    return xs_half + dt * v_new, v_new


def relativistic_rotation(dt, B, p_minus, q, m):
    """_particles.py:138-148."""
    c = speed_of_light
    B = np.asarray(B, dtype=np.float64)
    p_minus = np.asarray(p_minus, dtype=np.float64)
    gamma_minus = np.sqrt(1 + np.sum(p_minus ** 2) / (m ** 2 * c ** 2))
    t = (q * dt) / (2 * m * gamma_minus) * B
    p_dot_t = np.dot(p_minus, t)
    p_cross_t = np.cross(p_minus, t)
    t2 = np.dot(t, t)
    return (p_minus * (1 - t2) + 2 * (p_dot_t * t + p_cross_t)) / (1 + t2)


def boris_step_relativistic(dt, xs_half, vs, q_s, m_s, E_at_x, B_at_x):
    """Relativistic Boris push; q_s, m_s are (N,) (weight-scaled).  _particles.py:169-200."""
    c = speed_of_light
    q_s = np.asarray(q_s, dtype=np.float64).reshape(-1)
    m_s = np.asarray(m_s, dtype=np.float64).reshape(-1)
    xs_out = np.empty_like(np.asarray(xs_half, dtype=np.float64))
    vs_out = np.empty_like(xs_out)
    for p in range(len(xs_out)):
        v, q, m, E, B = vs[p], q_s[p], m_s[p], E_at_x[p], B_at_x[p]
        gamma_n = 1 / np.sqrt(1.0 - np.sum((v / c) ** 2))
        p_n = gamma_n * m * v
        p_minus = p_n + q * E * dt / 2
        p_plus = relativistic_rotation(dt, B, p_minus, q, m)
        p_new = p_plus + q * E * dt / 2
        gamma_new = np.sqrt(1.0 + np.sum((p_new / (m * c)) ** 2))
        v_new = p_new / (gamma_new * m)
        xs_out[p] = xs_half[p] + dt * v_new
        vs_out[p] = v_new
    return xs_out, vs_out


# ----------------------------------------------------------------------------------------------------
# deposition -- jaxincell/_sources.py:43-237
# ----------------------------------------------------------------------------------------------------
def charge_density_BCs(bc_left, bc_right, position, dx, grid, charge):
    """Spill-over of the S2 cloud past the end nodes, routed by the particle BCs.  _sources.py:60-81."""
    position = np.asarray(position, dtype=np.float64)
    extra_left = (charge / dx) * np.where(np.abs(position - grid[0]) <= dx / 2, 0.5 * (0.5 + (grid[0] - position) / dx) ** 2, 0.0)
    extra_right = (charge / dx) * np.where(np.abs(position - grid[-1]) <= dx / 2, 0.5 * (0.5 + (position - grid[-1]) / dx) ** 2, 0.0)
    zero = np.zeros_like(extra_left)
    to_left = {0: extra_right, 1: extra_left}.get(bc_left, zero)
    to_right = {0: extra_left, 1: extra_right}.get(bc_right, zero)
    return to_left, to_right


def single_particle_charge_density(x, q, dx, grid, bc_left, bc_right):
    """S2 cloud of particles on ALL G nodes: x, q scalars -> (G,) or arrays (N,) -> (N,G).  _sources.py:101-110."""
    x = np.asarray(x, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64)
    grid = np.asarray(grid, dtype=np.float64)
    sep = x[..., None] - grid
    r = np.abs(sep)
    core = 3 / 4 - sep ** 2 / (dx ** 2)
    wing = 0.5 * (3 / 2 - r / dx) ** 2
    rho = (q / dx)[..., None] * np.where(r <= dx / 2, core, np.where((dx / 2 < r) & (r <= 3 * dx / 2), wing, 0.0))
    to_left, to_right = charge_density_BCs(bc_left, bc_right, x, dx, grid, q)
    rho[..., 0] = to_left + rho[..., 0]
    rho[..., -1] = to_right + rho[..., -1]
    return rho


def calculate_charge_density(xs_n, qs, dx, grid, bc_left, bc_right, filter_passes=5, filter_alpha=0.5,
                             filter_strides=(1, 2, 4), field_BC_left=0, field_BC_right=0):
    """rho on the grid: sum of S2 clouds, then the digital filter.  _sources.py:136-154."""
    xs_n = np.asarray(xs_n, dtype=np.float64)
    qs = np.asarray(qs, dtype=np.float64)
    x = xs_n[:, 0] if xs_n.ndim == 2 else xs_n
    q = qs[:, 0] if qs.ndim == 2 else qs
    total = single_particle_charge_density(x, q, dx, grid, bc_left, bc_right).sum(axis=0)
    return filter_scalar_field(total, filter_passes, filter_alpha, filter_strides, field_BC_left, field_BC_right)


def current_density(xs_minus, xs_n, xs_plus, vs_n, qs, dx, dt, grid, grid_start, bc_left, bc_right,
                    filter_passes=5, filter_alpha=0.5, filter_strides=(1, 2, 4), field_BC_left=0, field_BC_right=0):
    """Charge-conserving J_x from a 6-node windowed prefix sum + J_y,z = rho(x_n) v.  _sources.py:185-237."""
    grid = np.asarray(grid, dtype=np.float64)
    G = len(grid)
    xs_minus = np.asarray(xs_minus, dtype=np.float64)
    xs_plus = np.asarray(xs_plus, dtype=np.float64)
    xs_n = np.asarray(xs_n, dtype=np.float64)
    vs_n = np.asarray(vs_n, dtype=np.float64)
    q = np.asarray(qs, dtype=np.float64).reshape(-1)
    J = np.zeros((G, 3))
    for p in range(len(q)):
        xm, xp = xs_minus[p, 0], xs_plus[p, 0]
        cell = int(np.floor_divide(xm - grid_start, dx))  # :190
        diff = (single_particle_charge_density(xp, q[p], dx, grid, bc_left, bc_right)
                - single_particle_charge_density(xm, q[p], dx, grid, bc_left, bc_right)) / dt  # :193-196
        short = np.roll(diff, 3 - cell)[:6]  # :199
        j_short = np.cumsum(-short * dx)  # :200
        jx = np.zeros(G)
        jx[:len(j_short)] = j_short  # :203-204
        jx = np.roll(jx, cell - 3)  # :207
        rho_n = single_particle_charge_density(xs_n[p, 0], q[p], dx, grid, bc_left, bc_right)  # :213
        J[:, 0] += jx
        J[:, 1] += rho_n * vs_n[p, 1]
        J[:, 2] += rho_n * vs_n[p, 2]
    return filter_vector_field(J, filter_passes, filter_alpha, filter_strides, field_BC_left, field_BC_right)


# ----------------------------------------------------------------------------------------------------
# digital filter -- jaxincell/_filters.py
# ----------------------------------------------------------------------------------------------------
_MAX_FILTER_PASSES = 16


def _shift_with_bc_1d(x, shift, bc_left, bc_right):
    """_filters.py:20-50.  Periodic only if BOTH ends are periodic; otherwise index-shift + clamp (+ zero for absorbing)."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    if bc_left == 0 and bc_right == 0:
        return np.roll(x, shift, axis=0)
    idx = np.arange(n) + shift
    out = x[np.clip(idx, 0, n - 1)].copy()
    mask = ((idx < 0) & (bc_left == 2)) | ((idx >= n) & (bc_right == 2))
    out[mask] = 0.0
    return out


def binomial_filter_3point(x, alpha=0.5, stride=1, bc_left=0, bc_right=0):
    """_filters.py:64-66."""
    x = np.asarray(x, dtype=np.float64)
    left = _shift_with_bc_1d(x, -stride, bc_left, bc_right)
    right = _shift_with_bc_1d(x, +stride, bc_left, bc_right)
    return alpha * x + (1 - alpha) * 0.5 * (left + right)


def _repeat_filter(y, stride, passes, alpha, bc_left=0, bc_right=0):
    """(passes-1) smoothing sweeps (capped at 16) + one compensation sweep.  _filters.py:85-122."""
    y = np.asarray(y, dtype=np.float64)
    if passes <= 0:
        return y
    passes_clamped = min(passes, _MAX_FILTER_PASSES + 1)
    num_regular = min(max(passes - 1, 0), _MAX_FILTER_PASSES)
    for _ in range(num_regular):
        y = binomial_filter_3point(y, alpha, stride, bc_left, bc_right)
    comp_alpha = passes_clamped - alpha * (passes_clamped - 1)
    return binomial_filter_3point(y, comp_alpha, stride, bc_left, bc_right)


def filter_scalar_field(f, passes=5, alpha=0.5, strides=(1, 2, 4), bc_left=0, bc_right=0):
    """_filters.py:146-153."""
    y = np.asarray(f, dtype=np.float64)
    for s in strides:
        y = _repeat_filter(y, int(s), passes, alpha, bc_left, bc_right)
    return y


def filter_vector_field(F, passes=5, alpha=0.5, strides=(1, 2, 4), bc_left=0, bc_right=0):
    """_filters.py:177-184 (filters along axis 0, every component alike)."""
    return filter_scalar_field(F, passes, alpha, strides, bc_left, bc_right)


# ----------------------------------------------------------------------------------------------------
# fields -- jaxincell/_fields.py:62-193
# ----------------------------------------------------------------------------------------------------
def E_from_Gauss_1D_Cartesian(charge_density, dx):
    """Dense lower-bidiagonal solve of the reference (== cumulative sum).  _fields.py:76-81."""
    rho = np.asarray(charge_density, dtype=np.float64)
    n = len(rho)
    D = np.diag(np.ones(n)) - np.diag(np.ones(n - 1), k=-1)
    return (dx / epsilon_0) * np.linalg.solve(D, rho)


def _gauss_wavenumbers(n, dx):
    kx = np.fft.fftfreq(n, d=dx) * 2 * np.pi
    kx[0] = 1.0  # "prevent division by zero" (_fields.py:26, :50)
    return kx


def E_from_Gauss_1D_FFT(charge_density, dx):
    """E_k = -i rho_k / (k eps0), real part of the inverse transform.  _fields.py:9-31."""
    rho = np.asarray(charge_density, dtype=np.float64)
    kx = _gauss_wavenumbers(len(rho), dx)
    E_k = -1j * np.fft.fft(rho) / kx / epsilon_0
    return np.fft.ifft(E_k).real


def E_from_Poisson_1D_FFT(charge_density, dx):
    """phi_k = -rho_k / (k^2 eps0) with phi_0 = 0, E_k = i k phi_k.  _fields.py:33-60."""
    rho = np.asarray(charge_density, dtype=np.float64)
    kx = _gauss_wavenumbers(len(rho), dx)
    phi_k = -np.fft.fft(rho) / kx ** 2 / epsilon_0
    phi_k[0] = 0.0
    return np.fft.ifft(1j * kx * phi_k).real


FIELD_SOLVERS = {1: E_from_Gauss_1D_FFT, 2: E_from_Gauss_1D_Cartesian, 3: E_from_Poisson_1D_FFT}  # _algorithms.py:73-77


def curlE(E, B, dx, dt, bc_left, bc_right):
    """Backward difference with a left ghost row.  _fields.py:102-111."""
    E = np.asarray(E, dtype=np.float64)
    gl, gr = field_ghost_cells_E(bc_left, bc_right, E, np.asarray(B, dtype=np.float64))
    P = np.concatenate([gl[None], E, gr[None]], axis=0)
    dFz = (P[1:-1, 2] - P[0:-2, 2]) / dx
    dFy = (P[1:-1, 1] - P[0:-2, 1]) / dx
    return np.stack([np.zeros(len(dFz)), -dFz, dFy], axis=1)


def curlB(B, E, dx, dt, bc_left, bc_right):
    """Forward difference with a right ghost row (note the roll by -1).  _fields.py:132-144."""
    B = np.asarray(B, dtype=np.float64)
    gl, gr = field_ghost_cells_B(bc_left, bc_right, B, np.asarray(E, dtype=np.float64))
    P = np.roll(np.concatenate([gl[None], B, gr[None]], axis=0), -1, axis=0)
    dFz = (P[1:-1, 2] - P[0:-2, 2]) / dx
    dFy = (P[1:-1, 1] - P[0:-2, 1]) / dx
    return np.stack([np.zeros(len(dFz)), -dFz, dFy], axis=1)


def field_update(E, B, dx, dt, j, bc_left, bc_right):
    """Simultaneous variant (not on the hot path; kept for the KAT).  _fields.py:164-173."""
    cE = curlE(E, B, dx, dt, bc_left, bc_right)
    cB = curlB(B, E, dx, dt, bc_left, bc_right)
    return E + dt * ((speed_of_light ** 2) * cB - (j / epsilon_0)), B - dt * cE


def field_update1(E, B, dx, dt, j, bc_left, bc_right):
    """E first (Ampere), then B (Faraday) with the new E.  _fields.py:178-183."""
    E = E + dt * ((speed_of_light ** 2) * curlB(B, E, dx, dt, bc_left, bc_right) - (j / epsilon_0))
    B = B - dt * curlE(E, B, dx, dt, bc_left, bc_right)
    return E, B


def field_update2(E, B, dx, dt, j, bc_left, bc_right):
    """B first (Faraday), then E (Ampere) with the new B.  _fields.py:188-193."""
    B = B - dt * curlE(E, B, dx, dt, bc_left, bc_right)
    E = E + dt * ((speed_of_light ** 2) * curlB(B, E, dx, dt, bc_left, bc_right) - (j / epsilon_0))
    return E, B


# ----------------------------------------------------------------------------------------------------
# the step and the driver -- jaxincell/_algorithms.py:17-95, jaxincell/_simulation.py:216-257
# ----------------------------------------------------------------------------------------------------
def Boris_step(carry, solver, ext_E, ext_B, dx, dt, grid, box_size, pbl, pbr, fbl, fbr):
    """One explicit step, BOTH current deposits included exactly as the reference orders them."""
    E, B, x_m, x_n, x_p, v, qs, ms, q_ms = carry
    fp, fa, fs = solver["filter_passes"], solver["filter_alpha"], solver["filter_strides"]
    gs = grid[0] - dx / 2
    J = current_density(x_m, x_n, x_p, v, qs, dx, dt, grid, gs, pbl, pbr, fp, fa, fs, fbl, fbr)  # :29-32
    E, B = field_update1(E, B, dx, dt / 2, J, fbl, fbr)  # :33
    tot_E = E + ext_E  # :36-37
    tot_B = B + ext_B
    n = len(qs)
    E_p = np.zeros((n, 3))
    B_p = np.zeros((n, 3))
    for p in range(n):  # :40-45
        E_p[p] = fields_to_particles_grid(x_p[p], tot_E, dx, grid + dx / 2, grid[0], fbl, fbr)
        B_p[p] = fields_to_particles_grid(x_p[p], tot_B, dx, grid, grid[0] - dx / 2, fbl, fbr)
    if solver.get("relativistic", False):  # :48-53
        x_pp, v_new = boris_step_relativistic(dt, x_p, v, qs, ms, E_p, B_p)
    else:
        x_pp, v_new = boris_step(dt, x_p, v, q_ms, E_p, B_p)
    x_pp, v_new, qs, ms, q_ms = set_BC_particles(x_pp, v_new, qs, ms, q_ms, dx, grid, *box_size, pbl, pbr)  # :56-58
    x_new = set_BC_positions(x_pp - (dt / 2) * v_new, dx, grid, *box_size, pbl, pbr)  # :60-61
    J = current_density(x_p, x_new, x_pp, v_new, qs, dx, dt, grid, gs, pbl, pbr, fp, fa, fs, fbl, fbr)  # :63-66
    E, B = field_update2(E, B, dx, dt / 2, J, fbl, fbr)  # :67
    field_solver = solver.get("field_solver", 0)
    if field_solver != 0:  # :69-78 -- `positions` is still x_n here (rebound at :84), `qs` already the post-BC charges of :56
        rho_faces = calculate_charge_density(x_n, qs, dx, grid + dx / 2, pbl, pbr, fp, fa, fs, fbl, fbr)
        E = E.copy()
        E[:, 0] = FIELD_SOLVERS[field_solver](rho_faces, dx)
    carry = (E, B, x_p, x_new, x_pp, v_new, qs, ms, q_ms)  # :81-87
    rho = calculate_charge_density(x_new, qs, dx, grid, pbl, pbr, fp, fa, fs, fbl, fbr)  # :90-92
    return carry, (x_new, v_new, E, B, J, rho)


def initial_fields(x0, qs, dx, grid, pbl, pbr, solver, fbl, fbr):
    """E_x from Gauss, everything else zero.  _state_initialization.py:371-378."""
    rho0 = calculate_charge_density(x0, qs, dx, grid, pbl, pbr, solver["filter_passes"], solver["filter_alpha"],
                                    solver["filter_strides"], fbl, fbr)
    G = len(grid)
    E = np.zeros((G, 3))
    E[:, 0] = E_from_Gauss_1D_Cartesian(rho0, dx)
    return E, np.zeros((G, 3))


def run(x0, v0, qs, ms, q_ms, *, length, G, dt, total_steps, box_yz=None, pbl=0, pbr=0, fbl=0, fbr=0,
        solver=None, ext_E=None, ext_B=None):
    """Leap-frog start-up + T Boris steps; returns the stacked histories the reference emits.

    _simulation.py:200-257.  x0, v0: (N,3); qs, ms, q_ms: (N,) weight-scaled charge / mass and q/m.
    """
    solver = {"filter_passes": 5, "filter_alpha": 0.5, "filter_strides": (1, 2, 4), "relativistic": False, **(solver or {})}
    dx = length / G
    grid = np.linspace(-length / 2 + dx / 2, length / 2 - dx / 2, G)  # _state_initialization.py:41
    Ly, Lz = box_yz if box_yz is not None else (length, length)
    box = (length, Ly, Lz)
    ext_E = np.zeros((G, 3)) if ext_E is None else np.asarray(ext_E, dtype=np.float32).astype(np.float64)
    ext_B = np.zeros((G, 3)) if ext_B is None else np.asarray(ext_B, dtype=np.float32).astype(np.float64)
    x0 = np.asarray(x0, dtype=np.float64)
    v0 = np.asarray(v0, dtype=np.float64)
    qs = np.asarray(qs, dtype=np.float64).reshape(-1)
    ms = np.asarray(ms, dtype=np.float64).reshape(-1)
    q_ms = np.asarray(q_ms, dtype=np.float64).reshape(-1)
    E, B = initial_fields(x0, qs, dx, grid, pbl, pbr, solver, fbl, fbr)
    # _simulation.py:217-225 -- note `velocities` is rebound to the post-BC value before the minus half-step.
    x_p, v, q1, m1, qm1 = set_BC_particles(x0 + (dt / 2) * v0, v0, qs, ms, q_ms, dx, grid, *box, pbl, pbr)
    x_m = set_BC_positions(x0 - (dt / 2) * v, dx, grid, *box, pbl, pbr)
    carry = (E, B, x_m, x0, x_p, v, q1, m1, qm1)
    hist = {k: [] for k in ("positions", "velocities", "electric_field", "magnetic_field", "current_density", "charge_density")}
    for _ in range(total_steps):
        carry, (xn, vn, En, Bn, Jn, rn) = Boris_step(carry, solver, ext_E, ext_B, dx, dt, grid, box, pbl, pbr, fbl, fbr)
        for k, a in zip(hist, (xn, vn, En, Bn, Jn, rn)):
            hist[k].append(np.array(a, copy=True))
    out = {k: np.stack(v_) for k, v_ in hist.items()}
    out.update(grid=grid, dx=dx, dt=dt, initial_velocities=v, fields=(E, B), final_carry=carry)
    return out


# ----------------------------------------------------------------------------------------------------
# implicit Crank-Nicolson stepper -- jaxincell/_algorithms.py:100-241, _sources.py:10-40,240-282, _particles.py:47-65
# (already O(1) per particle in the reference: vectorised here, not looped)
# ----------------------------------------------------------------------------------------------------
def get_S2_weights_and_indices_periodic_CN(x, dx, grid_start, grid_size):
    """Nearest node k = round((x - grid_start)/dx) (half to even, like jnp.round), nodes k-1,k,k+1 wrapped.  _sources.py:10-40."""
    x_norm = (np.asarray(x, dtype=np.float64) - grid_start) / dx
    k = np.round(x_norm).astype(np.int64)
    indices = np.stack([k - 1, k, k + 1], axis=-1) % grid_size
    d = x_norm - k
    weights = np.stack([0.5 * (0.5 - d) ** 2, 0.75 - d ** 2, 0.5 * (0.5 + d) ** 2], axis=-1)
    return indices, weights


def fields_to_particles_periodic_CN(xs, field, dx, grid_start):
    """_particles.py:47-65 for all particles: dot(weights, field[indices])."""
    idx, w = get_S2_weights_and_indices_periodic_CN(xs[:, 0], dx, grid_start, len(field))
    return np.einsum("pk,pkc->pc", w, np.asarray(field)[idx])


def current_density_periodic_CN(xs_n, vs_n, qs, dx, grid_start, grid_size):
    """J = (q/dx) v S(x) scattered with wrapped indices.  _sources.py:240-282."""
    idx, w = get_S2_weights_and_indices_periodic_CN(xs_n[:, 0], dx, grid_start, grid_size)
    J = np.zeros((grid_size, 3))
    for c in range(3):
        a = vs_n[:, c] * (qs / dx)
        J[:, c] = np.bincount(idx.ravel(), weights=(w * a[:, None]).ravel(), minlength=grid_size)
    return J


def CN_step(carry, solver, dx, dt, grid, box_size, pbl, pbr, fbl, fbr, num_substeps):
    """One implicit step: Picard iteration over (Faraday with E^{n+1/2}, sub-stepped particle push, Ampere with J - <J>).
    Returns (carry, step_data, n_iterations).  _algorithms.py:100-241."""
    E_field, B_field, positions, velocities, qs, ms, q_ms = carry
    E_start = grid[0] + dx / 2  # :110
    B_start = grid[0] - dx / 2  # :111
    G = len(grid)
    tol = solver["tolerance_Picard_iterations_implicit_CN"]
    max_iter = solver["max_number_of_Picard_iterations_implicit_CN"]
    dtau = dt / num_substeps
    c_sq = speed_of_light ** 2
    pos_stag = np.repeat(positions[None], num_substeps, axis=0)  # :121
    E_guess = E_field  # picard_init[1]
    # picard_init (:205): the values the loop returns when it never runs
    E_calc, B_next, pos_final, vel_final, J_iter = E_field, B_field, positions + dt * velocities, velocities, np.zeros_like(E_field)
    delta_E, i = np.inf, 0
    while delta_E > tol and i < max_iter:  # :213-215
        E_avg = 0.5 * (E_field + E_guess)  # :133
        B_next = B_field - dt * curlE(E_avg, B_field, dx, dt, fbl, fbr)  # :136-137
        B_avg = 0.5 * (B_next + B_field)  # :142
        pos_sub, vel_sub, q_sub, m_sub, qm_sub = positions, velocities, qs, ms, q_ms  # :185 (pos_fix, vel_fix, the outer charges)
        J_acc = np.zeros((G, 3))
        new_stag = np.empty_like(pos_stag)
        for s_ in range(num_substeps):  # :148-184
            stag_prev = pos_stag[s_]
            E_mid = fields_to_particles_periodic_CN(stag_prev, E_avg, dx, E_start)
            B_mid = fields_to_particles_periodic_CN(stag_prev, B_avg, dx, B_start)
            _, vel_new = boris_step(dtau, stag_prev, vel_sub, q_ms, E_mid, B_mid)  # :157 (the OUTER q/m)
            vel_mid = 0.5 * (vel_sub + vel_new)
            pos_new = pos_sub + vel_mid * dtau
            pos_new, vel_mid, q_new, m_new, qm_new = set_BC_particles(pos_new, vel_mid, q_sub, m_sub, qm_sub, dx, grid, *box_size, pbl, pbr)
            new_stag[s_] = set_BC_positions(pos_new - 0.5 * dtau * vel_mid, dx, grid, *box_size, pbl, pbr)
            J_acc += current_density_periodic_CN(stag_prev, vel_mid, qs, dx, E_start, G) * dtau  # :176-181 (the OUTER charges)
            pos_sub, vel_sub, q_sub, m_sub, qm_sub = pos_new, vel_new, q_new, m_new, qm_new
        pos_stag = new_stag
        pos_final, vel_final = pos_sub, vel_sub
        J_iter = J_acc / dt  # :190
        mean_J = J_iter.mean(axis=0)
        E_calc = E_field + dt * (c_sq * curlB(B_avg, E_field, dx, dt, fbl, fbr) - (1 / epsilon_0) * (J_iter - mean_J))  # :197-198
        delta_E = np.abs(np.max(E_calc - E_guess)) / (np.max(np.abs(E_calc)) + 1e-12)  # :224
        E_guess = E_calc
        i += 1
    rho = calculate_charge_density(pos_final, qs, dx, grid, pbl, pbr, 0, 0.5, (1, 2, 4), fbl, fbr)  # :236-238
    carry = (E_calc, B_next, pos_final, vel_final, qs, ms, q_ms)
    return carry, (pos_final, vel_final, E_calc, B_next, J_iter, rho), i


def run_CN(x0, v0, qs, ms, q_ms, *, length, G, dt, total_steps, box_yz=None, pbl=0, pbr=0, fbl=0, fbr=0, solver=None):
    """time_evolution_algorithm = 1: _simulation.py:216-257 with the CN carry (positions stay x0, velocities are the post-BC ones)."""
    solver = {"filter_passes": 5, "filter_alpha": 0.5, "filter_strides": (1, 2, 4), "max_number_of_Picard_iterations_implicit_CN": 20,
              "number_of_particle_substeps_implicit_CN": 2, "tolerance_Picard_iterations_implicit_CN": 1e-6, **(solver or {})}
    dx = length / G
    grid = np.linspace(-length / 2 + dx / 2, length / 2 - dx / 2, G)
    Ly, Lz = box_yz if box_yz is not None else (length, length)
    box = (length, Ly, Lz)
    x0 = np.asarray(x0, dtype=np.float64); v0 = np.asarray(v0, dtype=np.float64)
    qs = np.asarray(qs, dtype=np.float64).reshape(-1); ms = np.asarray(ms, dtype=np.float64).reshape(-1)
    q_ms = np.asarray(q_ms, dtype=np.float64).reshape(-1)
    E, B = initial_fields(x0, qs, dx, grid, pbl, pbr, solver, fbl, fbr)
    _, v, q1, m1, qm1 = set_BC_particles(x0 + (dt / 2) * v0, v0, qs, ms, q_ms, dx, grid, *box, pbl, pbr)
    carry = (E, B, x0, v, q1, m1, qm1)
    keys = ("positions", "velocities", "electric_field", "magnetic_field", "current_density", "charge_density")
    hist = {k: [] for k in keys}
    iters = []
    for _ in range(total_steps):
        carry, data, n_it = CN_step(carry, solver, dx, dt, grid, box, pbl, pbr, fbl, fbr, solver["number_of_particle_substeps_implicit_CN"])
        iters.append(n_it)
        for k, a in zip(keys, data):
            hist[k].append(np.array(a, copy=True))
    out = {k: np.stack(v_) for k, v_ in hist.items()}
    out.update(grid=grid, dx=dx, dt=dt, initial_velocities=v, fields=(E, B), picard_iterations=np.array(iters), final_carry=carry)
    return out
