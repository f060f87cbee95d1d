"""TEST INFRASTRUCTURE ONLY -- O(N)+O(G) closed-form NumPy restatement of the JAX-in-Cell Boris step.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline/reference legs may import this.

``oracle/literal.py`` follows the reference expression by expression (O(N*G)); this module states the same
arithmetic per particle on its 3-node / 6-node stencils so that it can run at 1e5..1e7
particles, and it is validated against the literal form in ``tests/test_oracle_closed_form.py``.
It also uses the two structural facts of SURVEY.md section 0:
  * the second current deposit of step n equals the first one of step n+1 (one deposit per step);
  * only x_{n+1/2} and v_n are carried between steps.
This is the algorithm the CUDA kernels implement, written in NumPy; parity status as in literal.py.

Citations are relative to the reference checkout.
"""
from __future__ import annotations

import numpy as np

from . import literal as lit
from .literal import epsilon_0, speed_of_light  # noqa: F401  (re-exported for tests)

PERIODIC, REFLECTIVE, ABSORBING = 0, 1, 2


class Domain:
    """Grid geometry of `_state_initialization.py:27-49` (dt is passed in, already CFL*dx/c)."""

    def __init__(self, length, G, dt, length_y=None, length_z=None):
        self.L = float(length)
        self.G = int(G)
        self.dx = self.L / self.G
        self.dt = float(dt)
        self.grid = np.linspace(-self.L / 2 + self.dx / 2, self.L / 2 - self.dx / 2, self.G)
        self.Ly = float(length_y) if length_y else self.L
        self.Lz = float(length_z) if length_z else self.L

    @property
    def box(self):
        return (self.L, self.Ly, self.Lz)


# ----------------------------------------------------------------------------------------------------
# S2 stencil of a particle (jaxincell/_sources.py:83-110 + :43-81), as (node, value) entries
# ----------------------------------------------------------------------------------------------------
def s2_entries(x, q, dom: Domain, pbl, pbr, faces=False):
    """Five (node, value) entries per particle: nodes c-1, c, c+1 (dropped if off-grid) + the two end nodes,
    which receive the folded ghost weight according to the particle BCs.

    The ghost weight exists only when the particle sits in the edge half cell (|x-g_0|<=dx/2, resp. g_{G-1}),
    exactly like `charge_density_BCs`; a particle outside the box simply loses its off-grid nodes.
    """
    G, dx, grid, L = dom.G, dom.dx, dom.grid, dom.L
    x = np.asarray(x, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64)
    if faces:
        # nodes g_k + dx/2 (`grid + dx/2` of _algorithms.py:70): the nearest face of a particle in the left half cell is
        # face -1, which is not on the grid -- its weight is dropped, exactly as single_particle_charge_density does
        grid = grid + dx / 2
        s = (x - grid[0]) / dx
        c = np.floor(s + 0.5).astype(np.int64)
        d = s - c
    else:
        s = (x - grid[0]) / dx
        c = np.floor(s + 0.5).astype(np.int64)
        inside = (x >= -L / 2) & (x <= L / 2)
        c = np.where(inside, np.clip(c, 0, G - 1), c)
        cc = np.clip(c, 0, G - 1)
        d = np.where(inside, (x - grid[cc]) / dx, s - c)
    w = np.stack([0.5 * (0.5 - d) ** 2, 0.75 - d ** 2, 0.5 * (0.5 + d) ** 2], axis=1)
    a = (q / dx)[:, None] * w
    nodes = c[:, None] + np.array([-1, 0, 1])[None, :]
    ok = (nodes >= 0) & (nodes < G)
    vals = np.where(ok, a, 0.0)
    nodes = np.where(ok, nodes, 0)
    # the ghost weight is switched by the reference's own comparison (|x-g_0| <= dx/2), so that a particle sitting
    # exactly on the wall follows whatever the rounding of `grid` decides there (_sources.py:60-69)
    ex_l = (q / dx) * np.where(np.abs(x - grid[0]) <= dx / 2, 0.5 * (0.5 + (grid[0] - x) / dx) ** 2, 0.0)
    ex_r = (q / dx) * np.where(np.abs(x - grid[-1]) <= dx / 2, 0.5 * (0.5 + (x - grid[-1]) / dx) ** 2, 0.0)
    to_first = ex_r * (pbl == PERIODIC) + ex_l * (pbl == REFLECTIVE)
    to_last = ex_l * (pbr == PERIODIC) + ex_r * (pbr == REFLECTIVE)
    nodes = np.concatenate([nodes, np.zeros((len(x), 1), np.int64), np.full((len(x), 1), G - 1, np.int64)], axis=1)
    vals = np.concatenate([vals, to_first[:, None], to_last[:, None]], axis=1)
    return nodes, vals


def deposit_rho_raw(x, q, dom, pbl, pbr, faces=False):
    """Unfiltered rho on cell centres (or on the faces g_k + dx/2): sum of S2 clouds (`_sources.py:136-142`)."""
    nodes, vals = s2_entries(x, q, dom, pbl, pbr, faces)
    return np.bincount(nodes.ravel(), weights=vals.ravel(), minlength=dom.G)


def gauss_kernel(G, dx):
    """The spectral solvers of `_fields.py:9-60` are one circulant operator: E = h (*) rho (circular convolution) with
    h[d] = 1/(G eps0) sum_{m != 0} sin(2 pi m d / G) / k_m, k_m = 2 pi fftfreq(G, dx)[m]; the k = 0 and Nyquist terms vanish
    in the real part.  This is what the CUDA path evaluates (k_gauss)."""
    m = np.arange(1, G)
    k = 2 * np.pi * np.fft.fftfreq(G, d=dx)[1:]
    d = np.arange(G)
    phase = (np.outer(d, m) % G) / G
    return (np.sin(2 * np.pi * phase) / k[None, :]).sum(axis=1) / (G * epsilon_0)


def solve_Ex(rho_faces, dx, field_solver):
    """E_x of `_algorithms.py:73-78` in the form the kernels use: circular convolution (1, 3) or prefix sum (2)."""
    G = len(rho_faces)
    if field_solver == 2:
        return (dx / epsilon_0) * np.cumsum(rho_faces)
    h = gauss_kernel(G, dx)
    idx = (np.arange(G)[:, None] - np.arange(G)[None, :]) % G
    return (h[idx] * rho_faces[None, :]).sum(axis=1)


def deposit_current_raw(x_old, x_mid, x_new, v_mid, q, dom, pbl, pbr):
    """Unfiltered J (G,3) of `_sources.py:185-225` in closed form.

    J_x: prefix sum of -(rho(x_new)-rho(x_old)) dx/dt over the window of min(6,G) nodes starting three nodes
    left of the cell of x_old (periodic roll whatever the BC); anything outside the window is dropped.
    J_y,z: rho(x_mid) * v_{y,z}.
    """
    G, dx, dt, grid = dom.G, dom.dx, dom.dt, dom.grid
    n = len(q)
    gs = grid[0] - dx / 2  # _algorithms.py:30
    cell = np.floor_divide(np.asarray(x_old, dtype=np.float64) - gs, dx).astype(np.int64)
    W = min(6, G)
    flat = np.zeros(n * W)
    rows = np.arange(n) * W
    for xx, sign in ((x_new, +1.0), (x_old, -1.0)):
        nodes, vals = s2_entries(xx, q, dom, pbl, pbr)
        rel = np.mod(nodes - (cell[:, None] - 3), G)
        hit = rel < W
        idx = (rows[:, None] + rel)[hit]
        flat += np.bincount(idx, weights=(sign * vals / dt)[hit], minlength=n * W)
    short = flat.reshape(n, W)
    jwin = np.cumsum(-short * dx, axis=1)
    knodes = np.mod(cell[:, None] - 3 + np.arange(W)[None, :], G)
    J = np.zeros((G, 3))
    J[:, 0] = np.bincount(knodes.ravel(), weights=jwin.ravel(), minlength=G)
    nodes, vals = s2_entries(x_mid, q, dom, pbl, pbr)
    v_mid = np.asarray(v_mid, dtype=np.float64)
    J[:, 1] = np.bincount(nodes.ravel(), weights=(vals * v_mid[:, 1:2]).ravel(), minlength=G)
    J[:, 2] = np.bincount(nodes.ravel(), weights=(vals * v_mid[:, 2:3]).ravel(), minlength=G)
    return J


# ----------------------------------------------------------------------------------------------------
# gather (jaxincell/_particles.py:29-45 called as in _algorithms.py:40-43)
# ----------------------------------------------------------------------------------------------------
def gather(x, field, dom, grid_nodes, grid_start, fbl, fbr):
    """Vectorised quadratic gather of a (G,3) field at positions x (N,), ghost rows by the field BCs."""
    dx = dom.dx
    L2, L1, R = lit.field_2_ghost_cells(fbl, fbr, field)
    padded = np.concatenate([L2[None], L1[None], field, R[None]], axis=0)
    nodes = np.concatenate([[grid_nodes[0] - dx], grid_nodes])
    i = np.floor_divide(x - grid_start + dx, dx).astype(np.int64)
    g = nodes[np.clip(i, 0, len(nodes) - 1)]
    n = len(padded)
    f0 = padded[np.clip(i, 0, n - 1)]
    f1 = padded[np.clip(i + 1, 0, n - 1)]
    f2 = padded[np.clip(i + 2, 0, n - 1)]
    t = ((g - x) / dx)[:, None]
    return 0.5 * f0 * (0.5 + t) ** 2 + f1 * (0.75 - ((g - x) ** 2 / dx ** 2)[:, None]) + 0.5 * f2 * (0.5 - t) ** 2


def gather_EB(x, tot_E, tot_B, dom, fbl, fbr):
    """E lives on faces g_k+dx/2, B on centres g_k (`_algorithms.py:41-42`)."""
    grid, dx = dom.grid, dom.dx
    return (gather(x, tot_E, dom, grid + dx / 2, grid[0], fbl, fbr),
            gather(x, tot_B, dom, grid, grid[0] - dx / 2, fbl, fbr))


# ----------------------------------------------------------------------------------------------------
# pushers (jaxincell/_particles.py:68-200), vectorised
# ----------------------------------------------------------------------------------------------------
def push_boris(dt, x, v, qm, E, B):
    qm = qm[:, None]
    vm = v + qm * E * dt / 2
    Rv = vm + 0.5 * dt * qm * np.cross(vm, B)
    Bv = 0.5 * qm * dt * B
    vp = (np.cross(Rv, Bv) + np.sum(Rv * Bv, axis=1, keepdims=True) * Bv + Rv) / (1 + np.sum(Bv * Bv, axis=1, keepdims=True))
    vn = vp + qm * E * dt / 2
    return x + dt * vn, vn


def push_boris_relativistic(dt, x, v, q, m, E, B):
    c = speed_of_light
    q = q[:, None]
    m = m[:, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        gamma_n = 1 / np.sqrt(1.0 - np.sum((v / c) ** 2, axis=1, keepdims=True))
        p_n = gamma_n * m * v
        p_minus = p_n + q * E * dt / 2
        gamma_minus = np.sqrt(1 + np.sum(p_minus ** 2, axis=1, keepdims=True) / (m ** 2 * c ** 2))
        t = (q * dt) / (2 * m * gamma_minus) * B
        pdt = np.sum(p_minus * t, axis=1, keepdims=True)
        pxt = np.cross(p_minus, t)
        t2 = np.sum(t * t, axis=1, keepdims=True)
        p_plus = (p_minus * (1 - t2) + 2 * (pdt * t + pxt)) / (1 + t2)
        p_new = p_plus + q * E * dt / 2
        gamma_new = np.sqrt(1.0 + np.sum((p_new / (m * c)) ** 2, axis=1, keepdims=True))
        vn = p_new / (gamma_new * m)
    return x + dt * vn, vn


# ----------------------------------------------------------------------------------------------------
# fused step
# ----------------------------------------------------------------------------------------------------
class State:
    """What the CUDA path carries between steps: fields, x_{n+1/2}, v_n, per-particle q/m/(q/m), and J^n."""


def start(x0, v0, qs, ms, q_ms, dom: Domain, pbl, pbr, fbl, fbr, solver, ext_E=None, ext_B=None):
    """Initial fields + leap-frog start-up + the first current deposit (K0 of the CUDA path).

    `_state_initialization.py:371-378`, `_simulation.py:217-225`, `_algorithms.py:29-32` (step 0).
    """
    G, dx, dt, grid = dom.G, dom.dx, dom.dt, dom.grid
    fp, fa, fs = solver["filter_passes"], solver["filter_alpha"], tuple(solver["filter_strides"])
    st = State()
    st.dom, st.bcs, st.solver = dom, (pbl, pbr, fbl, fbr), solver
    st.ext_E = np.zeros((G, 3)) if ext_E is None else np.asarray(ext_E, np.float32).astype(np.float64)
    st.ext_B = np.zeros((G, 3)) if ext_B is None else np.asarray(ext_B, np.float32).astype(np.float64)
    x0 = np.asarray(x0, np.float64)
    v0 = np.asarray(v0, np.float64)
    qs = np.asarray(qs, np.float64).reshape(-1)
    rho0 = lit.filter_scalar_field(deposit_rho_raw(x0[:, 0], qs, dom, pbl, pbr), fp, fa, fs, fbl, fbr)
    E = np.zeros((G, 3))
    E[:, 0] = (dx / epsilon_0) * np.cumsum(rho0)  # == the dense solve of _fields.py:76-81
    st.E, st.B = E, np.zeros((G, 3))
    st.E0, st.B0 = E.copy(), st.B.copy()
    x_p, v, q1, m1, qm1 = lit.set_BC_particles(x0 + (dt / 2) * v0, v0, qs, np.asarray(ms, np.float64).reshape(-1),
                                               np.asarray(q_ms, np.float64).reshape(-1), dx, grid, *dom.box, pbl, pbr)
    x_m = lit.set_BC_positions(x0 - (dt / 2) * v, dx, grid, *dom.box, pbl, pbr)
    st.x_half, st.v, st.q, st.m, st.qm = x_p, v, q1, m1, qm1
    st.x_n = x0.copy()
    st.initial_velocities = v.copy()
    Jraw = deposit_current_raw(x_m[:, 0], x0[:, 0], x_p[:, 0], v, q1, dom, pbl, pbr)
    st.J = lit.filter_vector_field(Jraw, fp, fa, fs, fbl, fbr)
    return st


def step(st):
    """One step: Maxwell half step (E,B) -> gather -> push -> BC -> one deposit -> Maxwell half step (B,E)."""
    dom = st.dom
    pbl, pbr, fbl, fbr = st.bcs
    dx, dt, grid = dom.dx, dom.dt, dom.grid
    fp, fa, fs = st.solver["filter_passes"], st.solver["filter_alpha"], tuple(st.solver["filter_strides"])
    E, B = lit.field_update1(st.E, st.B, dx, dt / 2, st.J, fbl, fbr)
    E_p, B_p = gather_EB(st.x_half[:, 0], E + st.ext_E, B + st.ext_B, dom, fbl, fbr)
    if st.solver.get("relativistic", False):
        x_pp, v_new = push_boris_relativistic(dt, st.x_half, st.v, st.q, st.m, E_p, B_p)
    else:
        x_pp, v_new = push_boris(dt, st.x_half, st.v, st.qm, E_p, B_p)
    x_pp, v_new, q, m, qm = lit.set_BC_particles(x_pp, v_new, st.q, st.m, st.qm, dx, grid, *dom.box, pbl, pbr)
    x_new = lit.set_BC_positions(x_pp - (dt / 2) * v_new, dx, grid, *dom.box, pbl, pbr)
    Jraw = deposit_current_raw(st.x_half[:, 0], x_new[:, 0], x_pp[:, 0], v_new, q, dom, pbl, pbr)
    J = lit.filter_vector_field(Jraw, fp, fa, fs, fbl, fbr)
    E, B = lit.field_update2(E, B, dx, dt / 2, J, fbl, fbr)
    field_solver = st.solver.get("field_solver", 0)
    if field_solver != 0:  # _algorithms.py:69-78: rho of x_n (start of the step) with the post-BC charges, on the faces
        rho_f = lit.filter_scalar_field(deposit_rho_raw(st.x_n[:, 0], q, dom, pbl, pbr, faces=True), fp, fa, fs, fbl, fbr)
        E = E.copy()
        E[:, 0] = solve_Ex(rho_f, dx, field_solver)
    rho = lit.filter_scalar_field(deposit_rho_raw(x_new[:, 0], q, dom, pbl, pbr), fp, fa, fs, fbl, fbr)
    st.E, st.B, st.J = E, B, J
    st.x_n = x_new
    st.x_half, st.v, st.q, st.m, st.qm = x_pp, v_new, q, m, qm
    return x_new, v_new, E, B, J, rho


def run(x0, v0, qs, ms, q_ms, *, length, G, dt, total_steps, box_yz=None, pbl=0, pbr=0, fbl=0, fbr=0,
        solver=None, ext_E=None, ext_B=None, keep_particles=True):
    """Same contract as ``literal.run`` (histories of x_{n+1}, v_{n+1}, E, B, J, rho)."""
    solver = {"filter_passes": 5, "filter_alpha": 0.5, "filter_strides": (1, 2, 4), "relativistic": False, **(solver or {})}
    Ly, Lz = box_yz if box_yz is not None else (None, None)
    dom = Domain(length, G, dt, Ly, Lz)
    st = start(x0, v0, qs, ms, q_ms, dom, pbl, pbr, fbl, fbr, solver, ext_E, ext_B)
    keys = ("positions", "velocities", "electric_field", "magnetic_field", "current_density", "charge_density")
    hist = {k: [] for k in keys}
    for _ in range(total_steps):
        res = step(st)
        for k, a in zip(keys, res):
            if not keep_particles and k in ("positions", "velocities"):
                continue
            hist[k].append(np.array(a, copy=True))
    out = {k: np.stack(v_) for k, v_ in hist.items() if v_}
    out.update(grid=dom.grid, dx=dom.dx, dt=dom.dt, initial_velocities=st.initial_velocities, fields=(st.E0, st.B0), state=st)
    return out


# ----------------------------------------------------------------------------------------------------
# diagnostics used as parity metrics (jaxincell/_diagnostics.py:98-146, examples/inference_two_stream.py:108-203)
# ----------------------------------------------------------------------------------------------------
def energies(out, masses, dx):
    """Field, kinetic and total energy histories (no external fields)."""
    e = (epsilon_0 / 2) * np.sum(np.sum(out["electric_field"] ** 2, axis=-1), axis=-1) * dx
    b = 1 / (2 * lit.mu_0) * np.sum(np.sum(out["magnetic_field"] ** 2, axis=-1), axis=-1) * dx
    res = {"electric_field_energy": e, "magnetic_field_energy": b}
    if "velocities" in out:
        ke = 0.5 * np.sum(np.asarray(masses).reshape(1, -1) * np.sum(out["velocities"] ** 2, axis=-1), axis=-1)
        res.update(kinetic_energy=ke, total_energy=e + b + ke)
    return res


def growth_rate(Ex_hist, dx, dt, total_steps, frac=(0.30, 0.50)):
    """Half the least-squares slope of ln(dx*sum E_x^2) over steps [0.30T, 0.50T) against time_array."""
    T = Ex_hist.shape[0]
    t = np.linspace(0, total_steps * dt, total_steps)[:T]
    y = np.log(dx * np.sum(Ex_hist ** 2, axis=1) + 1e-30)
    i0, i1 = int(frac[0] * T), int(frac[1] * T)
    A = np.stack([np.ones(i1 - i0), t[i0:i1] - t[i0]], axis=1)
    beta = np.linalg.solve(A.T @ A, A.T @ y[i0:i1])
    return 0.5 * beta[1]
