"""Pin the oracle against the known-answer vectors of the reference's own unit tests.

Each test names the reference test (file:line) whose hand-computed numbers it re-expresses in NumPy.
JAX is not available, so these vectors -- not a live run of the reference -- are what pins the oracle.
"""
import numpy as np
import pytest
from numpy.testing import assert_allclose

from oracle import literal as L

c = L.speed_of_light
eps0 = L.epsilon_0

FIELD4 = np.array([[0.0, 0.0, 0.0], [2.0, 10.0, -1.0], [5.0, 20.0, 3.0], [9.0, 30.0, 2.0]])


# ---------------------------------------------------------------- gather: tests/test_particles.py:29-159
def _g(x, bc, grid=None, gs=0.0, dx=1.0, field=FIELD4):
    grid = np.arange(4.0) if grid is None else grid
    return L.fields_to_particles_grid(np.array([x, 0.0, 0.0]), field, dx, grid, gs, bc, bc)


def test_gather_centre_stencil():
    assert_allclose(_g(1.0, 1), 0.125 * FIELD4[0] + 0.75 * FIELD4[1] + 0.125 * FIELD4[2])


@pytest.mark.parametrize("x,expect", [(-0.5, {0: 0.5 * FIELD4[-1] + 0.5 * FIELD4[0], 1: FIELD4[0], 2: 0.5 * FIELD4[0]}),
                                      (3.5, {0: 0.5 * FIELD4[-1] + 0.5 * FIELD4[0], 1: FIELD4[-1], 2: 0.5 * FIELD4[-1]})])
def test_gather_edges_all_bcs(x, expect):
    for bc, val in expect.items():
        assert_allclose(_g(x, bc), val)


def test_gather_staggering_and_dx():
    g = np.arange(4.0)
    assert_allclose(_g(1.0, 1, grid=g + 0.5, gs=g[0]), 0.5 * FIELD4[0] + 0.5 * FIELD4[1])  # E grid
    assert_allclose(_g(1.0, 1, grid=g, gs=g[0] - 0.5), 0.125 * FIELD4[0] + 0.75 * FIELD4[1] + 0.125 * FIELD4[2])  # B grid
    hg = 0.5 * np.arange(4.0)
    assert_allclose(_g(0.5, 1, grid=hg, gs=hg[0] - 0.25, dx=0.5), 0.125 * FIELD4[0] + 0.75 * FIELD4[1] + 0.125 * FIELD4[2])


# ---------------------------------------------------------------- rotation / Boris: tests/test_particles.py:219-325
def test_rotation_limits():
    v = np.array([1.0, 2.0, -0.5])
    Bz = np.array([0.0, 0.0, 2.0])
    assert_allclose(L.rotation(0.1, np.zeros(3), v, 1.0), v, atol=0)
    assert_allclose(L.rotation(0.1, Bz, v, 0.0), v, atol=0)
    rp = L.rotation(0.1, Bz, np.array([1.0, 0, 0]), 1.0)
    rn = L.rotation(0.1, Bz, np.array([1.0, 0, 0]), -1.0)
    assert_allclose(np.linalg.norm(rp), 1.0, rtol=1e-12)
    assert rp[1] < 0 < rn[1]
    assert_allclose(rp[0], rn[0])
    assert_allclose(rp[1], -rn[1])


def test_boris_step_limits():
    dt = 0.2
    x = np.array([[0.0, 1.0, 2.0], [1.0, -1.0, 0.5]])
    v = np.array([[1.0, 0.0, 0.5], [-0.5, 0.25, 1.0]])
    qm = np.array([[2.0], [-1.0]])
    Z = np.zeros_like(x)
    xz, vz = L.boris_step(dt, x, v, qm, Z, Z)
    assert_allclose(vz, v, atol=0)
    assert_allclose(xz, x + dt * v)
    E = np.array([[0.5, -1.0, 0.25], [1.0, 0.0, -0.5]])
    xe, ve = L.boris_step(dt, x, v, qm, E, Z)
    assert_allclose(ve, v + qm * E * dt, rtol=1e-12)
    assert_allclose(xe, x + dt * (v + qm * E * dt), rtol=1e-12)
    B = np.array([[0.0, 0.0, 2.0], [0.0, 1.5, 0.0]])
    xb, vb = L.boris_step(dt, x, v, qm, Z, B)
    assert_allclose(np.linalg.norm(vb, axis=1), np.linalg.norm(v, axis=1), rtol=1e-12)
    assert_allclose(xb, x + dt * vb)
    xc, vc = L.boris_step(dt, x, v, qm, E, B)
    vm = v + qm * E * dt / 2
    vr = np.stack([L.rotation(dt, B[i], vm[i], qm[i, 0]) for i in range(2)])
    assert_allclose(vc, vr + qm * E * dt / 2, rtol=1e-12)
    assert_allclose(xc, x + dt * vc, rtol=1e-12)


# ---------------------------------------------------------------- relativistic: tests/test_particles.py:328-461
def test_relativistic_rotation_and_push():
    Bz = np.array([0.0, 0.0, 2.0])
    p = np.array([1.0, 0.0, 0.25])
    assert_allclose(L.relativistic_rotation(0.1, np.zeros(3), p, 1.0, 1.0), p, atol=0)
    rp = L.relativistic_rotation(0.1, Bz, p, 1.0, 1.0)
    rn = L.relativistic_rotation(0.1, Bz, p, -1.0, 1.0)
    assert_allclose(np.linalg.norm(rp), np.linalg.norm(p), rtol=1e-12)
    assert rp[1] < 0 < rn[1]
    dt = 0.01
    x = np.array([[0.0, 1.0, 2.0], [1.0, -1.0, 0.5]])
    v = np.array([[0.05 * c, 0, 0], [0, -0.03 * c, 0.02 * c]])
    q = np.array([1.0, -2.0])
    m = np.array([1.0, 3.0])
    Z = np.zeros_like(x)
    xz, vz = L.boris_step_relativistic(dt, x, v, q, m, Z, Z)
    assert_allclose(vz, v, rtol=1e-9)
    assert_allclose(xz, x + dt * v, rtol=1e-9)
    E = np.array([[2e8, -1e8, 0.0], [0.0, 3e8, -2e8]])
    xe, ve = L.boris_step_relativistic(dt, x, v, q, m, E, Z)
    g0 = 1 / np.sqrt(1 - np.sum((v / c) ** 2, axis=1))
    p1 = g0[:, None] * m[:, None] * v + q[:, None] * E * dt
    g1 = np.sqrt(1 + np.sum((p1 / (m[:, None] * c)) ** 2, axis=1))
    v1 = p1 / (g1[:, None] * m[:, None])
    assert_allclose(ve, v1, rtol=1e-9)
    assert_allclose(xe, x + dt * v1, rtol=1e-9)
    assert np.all(np.linalg.norm(ve, axis=1) < c)
    B = np.array([[0.0, 0.0, 2.0], [0.0, -1.5, 0.5]])
    _, vb = L.boris_step_relativistic(dt, x, v, q, m, Z, B)
    assert_allclose(np.linalg.norm(vb, axis=1), np.linalg.norm(v, axis=1), rtol=1e-9)


# ---------------------------------------------------------------- deposition: tests/test_sources.py:78-328
@pytest.mark.parametrize("pos,bl,br,expected", [(0.25, 0, 0, (0.0, 0.0625)), (0.25, 1, 1, (0.0625, 0.0)), (0.25, 2, 2, (0.0, 0.0)),
                                                (2.75, 0, 0, (0.0625, 0.0)), (2.75, 1, 1, (0.0, 0.0625)), (2.75, 2, 2, (0.0, 0.0))])
def test_boundary_fold(pos, bl, br, expected):
    l, r = L.charge_density_BCs(bl, br, pos, 1.0, np.array([0.0, 1.0, 2.0, 3.0]), 2.0)
    assert_allclose([l, r], expected)


def test_s2_cloud_known_values():
    g = np.arange(5.0)
    assert_allclose(L.single_particle_charge_density(2.0, 2.0, 1.0, g, 2, 2), [0.0, 0.25, 1.5, 0.25, 0.0])
    assert_allclose(L.single_particle_charge_density(2.5, 2.0, 1.0, g, 2, 2), [0.0, 0.0, 1.0, 1.0, 0.0])
    assert_allclose(L.single_particle_charge_density(0.25, 2.0, 1.0, g, 0, 0), [1.375, 0.5625, 0.0, 0.0, 0.0625])
    assert_allclose(L.single_particle_charge_density(0.25, 2.0, 1.0, g, 1, 1), [1.4375, 0.5625, 0.0, 0.0, 0.0])
    assert_allclose(L.single_particle_charge_density(0.25, 2.0, 1.0, g, 2, 2), [1.375, 0.5625, 0.0, 0.0, 0.0])


def test_rho_sum_and_filter():
    g = np.arange(6.0)
    xs = np.array([[2.0], [3.25]])
    qs = np.array([[2.0], [-1.0]])
    manual = sum(L.single_particle_charge_density(xs[i, 0], qs[i, 0], 1.0, g, 0, 0) for i in range(2))
    assert_allclose(L.calculate_charge_density(xs, qs, 1.0, g, 0, 0, 0, 0.5, (1,)), manual)
    assert_allclose(L.calculate_charge_density(np.array([[2.0], [2.0]]), np.array([[2.0], [-2.0]]), 1.0, g, 0, 0, 0, 0.5, (1,)), 0 * g)
    f = L.calculate_charge_density(xs, qs, 1.0, g, 0, 0, 2, 0.4, (1,), 1, 1)
    assert_allclose(f, L.filter_scalar_field(manual, 2, 0.4, (1,), 1, 1))


def test_current_continuity_and_transverse():
    dx, dt = 1.0, 0.2
    g = np.arange(8.0)
    xm, xn, xp = np.array([[3.0]]), np.array([[3.1]]), np.array([[3.2]])
    v = np.array([[0.0, 0.5, -0.25]])
    q = np.array([[2.0]])
    J = L.current_density(xm, xn, xp, v, q, dx, dt, g, 0.0, 0, 0, 0, 0.5, (1,), 0, 0)
    rm = L.single_particle_charge_density(3.0, 2.0, dx, g, 0, 0)
    rp = L.single_particle_charge_density(3.2, 2.0, dx, g, 0, 0)
    assert_allclose((rp - rm) / dt + (J[:, 0] - np.roll(J[:, 0], 1)) / dx, 0 * g, atol=1e-12)
    rn = L.single_particle_charge_density(3.1, 2.0, dx, g, 0, 0)
    assert_allclose(J[:, 1], rn * 0.5)
    assert_allclose(J[:, 2], rn * -0.25)
    Jf = L.current_density(xm, xn, xp, v, q, dx, dt, g, 0.0, 1, 1, 2, 0.4, (1,), 1, 1)
    Ju = L.current_density(xm, xn, xp, v, q, dx, dt, g, 0.0, 1, 1, 0, 0.4, (1,), 1, 1)
    assert_allclose(Jf, L.filter_vector_field(Ju, 2, 0.4, (1,), 1, 1))


# ---------------------------------------------------------------- particle BCs: tests/test_boundary_conditions.py:177-387
def _bc1(x, v, bl, br):
    g = np.linspace(-1.0, 1.0, 10)
    xs, vs, q, m, qm = L.set_BC_particles(np.array([x]), np.array([v]), np.array([1.0]), np.array([1.0]), np.array([1.0]),
                                          0.1, g, 2.0, 2.0, 2.0, bl, br)
    return xs[0], vs[0], q[0], qm[0], g


def test_particle_bc_single():
    x, v, q, qm, _ = _bc1([1.0, 1.0, 1.0], [1.0, 1.0, 1.0], 0, 0)
    assert_allclose(x, [1.0, -1.0, -1.0]); assert_allclose(v, [1, 1, 1]); assert q == 1 and qm == 1
    x, v, q, qm, _ = _bc1([-1.1, 1.0, 1.0], [1.0, 1.0, 1.0], 1, 1)
    assert_allclose(x, [-0.9, -1.0, -1.0]); assert_allclose(v, [-1, 1, 1]); assert q == 1 and qm == 1
    x, v, q, qm, _ = _bc1([1.1, 1.0, 1.0], [1.0, 1.0, 1.0], 2, 2)
    assert_allclose(x, [1.3, -1.0, -1.0]); assert_allclose(v, [0, 0, 0]); assert q == 0 and qm == 0
    x, v, q, qm, _ = _bc1([-1.1, 1.0, 1.0], [1.0, 1.0, 1.0], 1, 2)
    assert_allclose(x, [-0.9, -1.0, -1.0]); assert_allclose(v, [-1, 1, 1]); assert q == 1 and qm == 1


def test_particle_bc_batched():
    g = np.linspace(-5.0, 5.0, 100)
    dx = 0.1
    q = np.array([1.0, -1.0]); m = np.array([1.0, 1.0])
    xs = np.array([[1.0, 2.0, 3.0], [-1.0, -2.0, -3.0]]); vs = np.array([[0.1, 0.2, 0.3], [-0.1, -0.2, -0.3]])
    o = L.set_BC_particles(xs, vs, q, m, q, dx, g, 10.0, 10.0, 10.0, 0, 0)
    assert_allclose(o[0], xs); assert_allclose(o[1], vs); assert_allclose(o[2], q); assert_allclose(o[4], q)
    xs = np.array([[6.0, 2.0, 3.0], [-6.0, -2.0, -3.0]])
    o = L.set_BC_particles(xs, vs, q, m, q, dx, g, 10.0, 10.0, 10.0, 1, 1)
    assert_allclose(o[0], [[4.0, 2.0, 3.0], [-4.0, -2.0, -3.0]]); assert_allclose(o[1], [[-0.1, 0.2, 0.3], [0.1, -0.2, -0.3]])
    assert_allclose(o[2], q)
    o = L.set_BC_particles(xs, vs, q, m, q, dx, g, 10.0, 10.0, 10.0, 2, 2)
    assert_allclose(o[0], [[g[-1] + 3 * dx, 2.0, 3.0], [g[0] - 1.5 * dx, -2.0, -3.0]])
    assert_allclose(o[1], 0 * vs); assert_allclose(o[2], [0, 0]); assert_allclose(o[3], m); assert_allclose(o[4], [0, 0])


def test_positions_only_bc():
    g = np.linspace(-1.0, 1.0, 10)
    f = lambda x, bl, br: L.set_BC_positions(np.array([x]), 0.1, g, 2.0, 2.0, 2.0, bl, br)[0]
    assert_allclose(f([1.0, 1.0, 1.0], 0, 0), [1.0, -1.0, -1.0])
    assert_allclose(f([-1.1, 1.0, 1.0], 1, 1), [-0.9, -1.0, -1.0])
    assert_allclose(f([-1.1, 1.0, 1.0], 2, 2), [g[0] - 0.15, -1.0, -1.0])


# ---------------------------------------------------------------- ghost cells: tests/test_boundary_conditions.py:389-459
def test_field_ghost_cells():
    E = np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]]); B = np.array([[0.1, 0.2, 0.3], [0.4, 0.5, 0.6]])
    l, r = L.field_ghost_cells_E(0, 0, E, B); assert_allclose(l, E[-1]); assert_allclose(r, E[0])
    l, r = L.field_ghost_cells_E(1, 1, E, B); assert_allclose(l, E[0]); assert_allclose(r, E[-1])
    l, r = L.field_ghost_cells_E(2, 2, E, B)
    assert_allclose(l, [0, -2 * c * B[0, 2] - E[0, 1], 2 * c * B[0, 1] - E[0, 2]])
    assert_allclose(r, [0, 3 * E[-1, 1] - 2 * c * B[-1, 2], 3 * E[-1, 2] + 2 * c * B[-1, 1]])
    Bf, Ef = E, B
    l, r = L.field_ghost_cells_B(0, 0, Bf, Ef); assert_allclose(l, Bf[-1]); assert_allclose(r, Bf[0])
    l, r = L.field_ghost_cells_B(1, 1, Bf, Ef); assert_allclose(l, Bf[0]); assert_allclose(r, Bf[-1])
    l, r = L.field_ghost_cells_B(2, 2, Bf, Ef)
    assert_allclose(l, [0, 3 * Bf[0, 1] - (2 / c) * Ef[0, 2], 3 * Bf[0, 2] + (2 / c) * Ef[0, 1]])
    assert_allclose(r, [0, -(2 / c) * Ef[-1, 2] - Bf[-1, 1], (2 / c) * Ef[-1, 1] - Bf[-1, 2]])
    l2, l1, r = L.field_2_ghost_cells(0, 0, E); assert_allclose(l2, E[-2]); assert_allclose(l1, E[-1]); assert_allclose(r, E[0])
    l2, l1, r = L.field_2_ghost_cells(1, 1, E); assert_allclose(l2, E[1]); assert_allclose(l1, E[0]); assert_allclose(r, E[-1])
    l2, l1, r = L.field_2_ghost_cells(2, 2, E); assert_allclose([l2, l1, r], 0)


# ---------------------------------------------------------------- fields: tests/test_fields.py:122-507
def test_gauss_cumsum():
    assert_allclose(L.E_from_Gauss_1D_Cartesian(eps0 * np.array([1.0, -0.25, 0.5, -1.25]), 0.5), [0.5, 0.375, 0.625, 0.0], atol=1e-15)


LIN = np.array([[0.0, 1.0, 2.0], [0.0, 3.0, 5.0], [0.0, 7.0, 11.0]])
Z3 = np.zeros((3, 3))


def test_curlE_vectors():
    const = np.tile([0.0, 3.0, -2.0], (3, 1))
    assert_allclose(L.curlE(const, Z3, 1.0, 0.1, 1, 1), Z3, atol=0)
    assert_allclose(L.curlE(LIN, Z3, 1.0, 0.1, 1, 1), [[0, 0, 0], [0, -3, 2], [0, -6, 4]])
    assert_allclose(L.curlE(LIN, Z3, 1.0, 0.1, 0, 0), [[0, 9, -6], [0, -3, 2], [0, -6, 4]])
    assert_allclose(L.curlE(LIN, Z3, 1.0, 0.1, 2, 2), [[0, -4, 2], [0, -3, 2], [0, -6, 4]])
    Bc = Z3.copy(); Bc[0] = [0.0, 1e-9, -2e-9]
    gl = np.array([0.0, -2 * c * Bc[0, 2] - LIN[0, 1], 2 * c * Bc[0, 1] - LIN[0, 2]])
    exp = np.array([[0, -(LIN[0, 2] - gl[2]) / 0.5, (LIN[0, 1] - gl[1]) / 0.5],
                    [0, -(LIN[1, 2] - LIN[0, 2]) / 0.5, (LIN[1, 1] - LIN[0, 1]) / 0.5],
                    [0, -(LIN[2, 2] - LIN[1, 2]) / 0.5, (LIN[2, 1] - LIN[1, 1]) / 0.5]])
    assert_allclose(L.curlE(LIN, Bc, 0.5, 0.1, 2, 1), exp, rtol=1e-12)


def test_curlB_vectors():
    const = np.tile([0.0, 3.0, -2.0], (3, 1))
    assert_allclose(L.curlB(const, Z3, 1.0, 0.1, 1, 1), Z3, atol=0)
    assert_allclose(L.curlB(LIN, Z3, 1.0, 0.1, 1, 1), [[0, -3, 2], [0, -6, 4], [0, 0, 0]])
    assert_allclose(L.curlB(LIN, Z3, 1.0, 0.1, 0, 0), [[0, -3, 2], [0, -6, 4], [0, 9, -6]])
    assert_allclose(L.curlB(LIN, Z3, 1.0, 0.1, 2, 2), [[0, -3, 2], [0, -6, 4], [0, 22, -14]])
    Ec = Z3.copy(); Ec[-1] = [0.0, 1.5e6, -0.5e6]
    gr = np.array([0.0, -(2 / c) * Ec[-1, 2] - LIN[-1, 1], (2 / c) * Ec[-1, 1] - LIN[-1, 2]])
    exp = np.array([[0, -(LIN[1, 2] - LIN[0, 2]) / 0.5, (LIN[1, 1] - LIN[0, 1]) / 0.5],
                    [0, -(LIN[2, 2] - LIN[1, 2]) / 0.5, (LIN[2, 1] - LIN[1, 1]) / 0.5],
                    [0, -(gr[2] - LIN[2, 2]) / 0.5, (gr[1] - LIN[2, 1]) / 0.5]])
    assert_allclose(L.curlB(LIN, Ec, 0.5, 0.1, 1, 2), exp, rtol=1e-12)


@pytest.mark.parametrize("upd", [L.field_update, L.field_update1, L.field_update2])
def test_field_update_current_response(upd):
    Z = np.zeros((4, 3))
    E, B = upd(Z, Z, 1.0, 0.2, Z, 1, 1)
    assert_allclose(E, Z, atol=0); assert_allclose(B, Z, atol=0)
    J = eps0 * np.tile([1.0, -2.0, 3.0], (4, 1))
    E, B = upd(Z, Z, 1.0, 0.2, J, 1, 1)
    assert_allclose(E, -0.2 * J / eps0); assert_allclose(B, Z, atol=0)


def test_field_update_ordering():
    dx, dt, bl, br = 0.5, 1e-6, 2, 1
    E = 1e-3 * np.array([[0, 1.0, 2.0], [0, -1.5, 0.5], [0, 0.25, -0.75], [0, 2.0, -1.0]])
    B = 1e-9 * np.array([[0, -2.0, 1.0], [0, 0.5, 3.0], [0, 1.5, -1.0], [0, -0.25, 2.0]])
    J = eps0 * np.array([[0, 0.5, -1.0], [0, -0.25, 0.75], [0, 1.25, -0.5], [0, -0.75, 0.25]])
    cE0 = L.curlE(E, B, dx, dt, bl, br); cB0 = L.curlB(B, E, dx, dt, bl, br)
    E1, B1 = L.field_update1(E, B, dx, dt, J, bl, br)
    eE = E + dt * (c ** 2 * cB0 - J / eps0)
    assert_allclose(E1, eE, rtol=1e-12); assert_allclose(B1, B - dt * L.curlE(eE, B, dx, dt, bl, br), rtol=1e-12)
    E2, B2 = L.field_update2(E, B, dx, dt, J, bl, br)
    eB = B - dt * cE0
    assert_allclose(B2, eB, rtol=1e-12)
    assert_allclose(E2, E + dt * (c ** 2 * L.curlB(eB, E, dx, dt, bl, br) - J / eps0), rtol=1e-12)
    Es, Bs = L.field_update(E, B, dx, dt, J, bl, br)
    assert not np.array_equal(B1, Bs) and not np.array_equal(E2, Es)


# ---------------------------------------------------------------- filter: tests/test_filters.py:18-109,274-336,416-472
def test_filter_identity_and_manual():
    x = np.linspace(0.0, 1.0, 10)
    assert_allclose(L.binomial_filter_3point(x, 1.0, 1), x, atol=1e-12)
    assert_allclose(L.binomial_filter_3point(x, 1.0, 2), x, atol=1e-12)
    x = np.arange(5.0)
    assert_allclose(L.binomial_filter_3point(x, 0.5, 1), 0.5 * x + 0.25 * (np.roll(x, 1) + np.roll(x, -1)), atol=1e-12)


def test_repeat_filter_semantics():
    x = np.linspace(-1.0, 1.0, 11)
    for p in (0, -3, 1):
        assert_allclose(L._repeat_filter(x, 1, p, 0.3), x, atol=1e-12)
    x = np.cos(np.linspace(0.0, 2 * np.pi, 17))
    for p in (2, 3, 5):
        y = x.copy()
        for _ in range(p - 1):
            y = L.binomial_filter_3point(y, 0.4, 1)
        y = L.binomial_filter_3point(y, p - 0.4 * (p - 1), 1)
        assert_allclose(L._repeat_filter(x, 1, p, 0.4), y, rtol=1e-10, atol=1e-12)
    big = L._repeat_filter(np.sin(np.linspace(0, 4 * np.pi, 33)), 1, 40, 0.5)
    assert big.shape == (33,) and np.all(np.isfinite(big))


def test_filter_nonperiodic_edges():
    x = np.arange(5.0)
    ref = [0.5 * x[j] + 0.25 * (x[max(j - 1, 0)] + x[min(j + 1, 4)]) for j in range(5)]
    assert_allclose(L.binomial_filter_3point(x, 0.5, 1, 1, 1), ref, atol=1e-12)
    x = np.array([1.0, 2.0, 3.0, 4.0])
    ref = [0.5 * x[j] + 0.25 * ((x[j - 1] if j >= 1 else 0.0) + (x[j + 1] if j + 1 < 4 else 0.0)) for j in range(4)]
    assert_allclose(L.binomial_filter_3point(x, 0.5, 1, 2, 2), ref, atol=1e-12)


def test_shift_semantics():
    x = np.arange(5.0)
    for s in (-2, -1, 1, 2):
        assert_allclose(L._shift_with_bc_1d(x, s, 0, 0), np.roll(x, s))
    assert_allclose(L._shift_with_bc_1d(x, 1, 1, 1), [1, 2, 3, 4, 4])
    assert_allclose(L._shift_with_bc_1d(x, -1, 1, 1), [0, 0, 1, 2, 3])
    assert_allclose(L._shift_with_bc_1d(x, 2, 1, 1), [2, 3, 4, 4, 4])
    assert_allclose(L._shift_with_bc_1d(x, 1, 2, 2), [1, 2, 3, 4, 0])
    assert_allclose(L._shift_with_bc_1d(x, -1, 2, 2), [0, 0, 1, 2, 3])
    assert_allclose(L._shift_with_bc_1d(x, 2, 2, 2), [2, 3, 4, 0, 0])


def test_constants():
    # tests/test_constants.py -- values are part of the results
    assert (L.epsilon_0, L.mu_0, L.speed_of_light) == (8.85418782e-12, 1.25663706e-6, 2.99792458e8)
    assert (L.elementary_charge, L.mass_electron, L.mass_proton) == (1.60217663e-19, 9.10938371e-31, 1.67262193e-27)


# ---------------------------------------------------------------------------------------------------------------------
# spectral field solvers -- reference tests/test_fields.py:28-120
# ---------------------------------------------------------------------------------------------------------------------
def test_E_from_Gauss_1D_FFT_modes():
    G, dx = 16, 0.25
    x = dx * np.arange(G)
    k = 2 * np.pi * 2 / (G * dx)
    assert np.all(L.E_from_Gauss_1D_FFT(np.zeros(G), dx) == 0)
    np.testing.assert_allclose(L.E_from_Gauss_1D_FFT(L.epsilon_0 * np.ones(G), dx), 0, atol=1e-12)  # k = 0 leaks no DC field
    np.testing.assert_allclose(L.E_from_Gauss_1D_FFT(L.epsilon_0 * np.sin(k * x), dx), -np.cos(k * x) / k, rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(L.E_from_Gauss_1D_FFT(3.5 * L.epsilon_0 * np.cos(k * x), dx), 3.5 * np.sin(k * x) / k, rtol=1e-6, atol=1e-12)


def test_E_from_Poisson_1D_FFT_modes():
    G, dx = 16, 0.25
    x = dx * np.arange(G)
    k = 2 * np.pi * 2 / (G * dx)
    assert np.all(L.E_from_Poisson_1D_FFT(np.zeros(G), dx) == 0)
    for rho, want in ((L.epsilon_0 * np.sin(k * x), -np.cos(k * x) / k), (3.5 * L.epsilon_0 * np.cos(k * x), 3.5 * np.sin(k * x) / k),
                      (L.epsilon_0 * (7.0 + np.sin(k * x)), -np.cos(k * x) / k)):
        np.testing.assert_allclose(L.E_from_Poisson_1D_FFT(rho, dx), want, rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(L.E_from_Poisson_1D_FFT(L.epsilon_0 * np.sin(k * x), dx), L.E_from_Gauss_1D_FFT(L.epsilon_0 * np.sin(k * x), dx),
                               rtol=1e-6, atol=1e-12)


@pytest.mark.parametrize("G,dx", [(16, 0.25), (17, 0.1), (70, 1e-4), (128, 3e-3)])
def test_circulant_form_of_the_spectral_solvers(G, dx):
    """What the CUDA path evaluates (closed_form.solve_Ex: circular convolution with gauss_kernel) is the FFT solve."""
    from oracle import closed_form as C
    rho = np.random.default_rng(G).standard_normal(G)
    for fs in (1, 2, 3):
        a, b = L.FIELD_SOLVERS[fs](rho, dx), C.solve_Ex(rho, dx, fs)
        np.testing.assert_allclose(b, a, rtol=0, atol=1e-13 * np.abs(a).max())


# ---------------------------------------------------------------------------------------------------------------------
# periodic S2 helpers of the Crank-Nicolson stepper -- reference tests/test_sources.py:25-76,331-372, tests/test_particles.py:162-218
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("x,weights", [(0.0, [0.125, 0.75, 0.125]), (0.5, [0.0, 0.5, 0.5]), (-0.25, [0.28125, 0.6875, 0.03125]),
                                       (4.9, [0.18, 0.74, 0.08])])
def test_S2_weights_and_indices_periodic_CN(x, weights):
    idx, w = L.get_S2_weights_and_indices_periodic_CN(np.array([x]), 1.0, 0.0, 5)
    assert idx[0].tolist() == [4, 0, 1]
    assert_allclose(w[0], weights, rtol=1e-12, atol=1e-15)
    assert_allclose(w.sum(), 1.0)


def test_current_density_periodic_CN_accumulates_wrapped_particles():
    J = L.current_density_periodic_CN(np.array([[0.0, 0, 0]]), np.array([[1.0, 2.0, -3.0]]), np.array([2.0]), 1.0, 0.0, 5)
    want = np.zeros((5, 3)); want[4] = [0.25, 0.5, -0.75]; want[0] = [1.5, 3.0, -4.5]; want[1] = [0.25, 0.5, -0.75]
    assert_allclose(J, want)
    J = L.current_density_periodic_CN(np.array([[0.0, 0, 0], [5.0, 0, 0]]), np.array([[1.0, 2.0, -3.0], [4.0, -2.0, 0.5]]), np.array([2.0, 1.0]),
                                      1.0, 0.0, 5)
    tot = np.array([2.0, 4.0, -6.0]) + np.array([4.0, -2.0, 0.5])
    want = np.zeros((5, 3)); want[4] = 0.125 * tot; want[0] = 0.75 * tot; want[1] = 0.125 * tot
    assert_allclose(J, want)


def test_fields_to_particles_periodic_CN():
    field = np.array([[0.0, 0.0, 0.0], [1.0, 10.0, -1.0], [2.0, 20.0, -2.0], [3.0, 30.0, -3.0], [4.0, 40.0, -4.0]])
    for x in (0.0, -0.25, 4.9, 5.0):
        idx, w = L.get_S2_weights_and_indices_periodic_CN(np.array([x]), 1.0, 0.0, 5)
        assert_allclose(L.fields_to_particles_periodic_CN(np.array([[x, 0.0, 0.0]]), field, 1.0, 0.0)[0], w[0] @ field[idx[0]], rtol=1e-12)
    const = np.tile(np.array([2.0, -3.0, 5.0]), (5, 1))
    assert_allclose(L.fields_to_particles_periodic_CN(np.array([[4.75, 0.0, 0.0]]), const, 1.0, 0.0)[0], const[0], rtol=1e-12)
    half = L.fields_to_particles_periodic_CN(np.array([[2.0 + 0.25, 0.0, 0.0]]), field, 0.5, 2.0)[0]
    assert_allclose(half, 0.5 * field[0] + 0.5 * field[1], rtol=1e-12)


def test_CN_step_picard_stopping_conditions_and_energy():
    """reference tests/test_algorithms.py:616-680 (stopping rules) + what the scheme is for: energy conservation."""
    from plasma import cfl_dt, two_species
    G, length = 16, 0.01
    p = two_species(300, 300, length=length, G=G, seed=3, vth_e=0.05, vth_yz=0.02, drift=3e7, plus_minus=True, gpdl=0.02)
    dt = cfl_dt(length, G, 0.3)
    kw = dict(length=length, G=G, dt=dt, total_steps=4)
    run = lambda **s: L.run_CN(p["x0"], p["v0"], p["q"], p["m"], p["qm"], solver=s, **kw)
    assert run(tolerance_Picard_iterations_implicit_CN=1e9, max_number_of_Picard_iterations_implicit_CN=5)["picard_iterations"].tolist() == [1] * 4
    assert run(tolerance_Picard_iterations_implicit_CN=1e-30, max_number_of_Picard_iterations_implicit_CN=3)["picard_iterations"].tolist() == [3] * 4
    out = run(tolerance_Picard_iterations_implicit_CN=1e-10, max_number_of_Picard_iterations_implicit_CN=30, number_of_particle_substeps_implicit_CN=3)
    assert 1 < out["picard_iterations"].max() < 30
    from oracle import closed_form as C
    e = C.energies(out, p["m"], out["dx"])["total_energy"]
    assert abs(e[-1] / e[0] - 1) < 1e-6
