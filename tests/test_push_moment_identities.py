"""The algebra behind the binned push kernel's deposition (header of jax-in-cell_b200/csrc/jic_push.cuh), checked on the CPU.

k_push never evaluates per-node S2 weights: for an offset t in (-3/2, 3/2) from the bin's node c it accumulates sums of
1, t, t^2, P(t) = max(t - 1/2, 0)^2 and N(t) = max(-t - 1/2, 0)^2 and forms the node values once per work item from the identities
below (truncated-power form of the quadratic B-spline of jaxincell/_sources.py:83-110).  These tests state the identities exactly as
the header does and compare them with the spline itself, for the centred weights, the cumulative weights of the charge-conserving J_x
(_sources.py:190-207) and the face weights of the field_solver deposit (_algorithms.py:69-72)."""
import numpy as np
import pytest


def s2(u):
    """The S2 shape of _sources.py:101-104 in units of dx."""
    a = np.abs(u)
    return np.where(a <= 0.5, 0.75 - u ** 2, np.where(a <= 1.5, 0.5 * (1.5 - a) ** 2, 0.0))


T = np.concatenate([np.linspace(-1.4999, 1.4999, 4001), [-1.0, -0.5, 0.0, 0.5, 1.0]])


def test_node_weights_from_moments():
    t = T
    P, N = np.maximum(t - 0.5, 0) ** 2, np.maximum(-t - 0.5, 0) ** 2
    w = {-2: N / 2, -1: ((t - 0.5) ** 2 - P - 3 * N) / 2, 0: 0.75 - t ** 2 + 1.5 * (N + P), 1: ((t + 0.5) ** 2 - N - 3 * P) / 2, 2: P / 2}
    for k, wk in w.items():
        np.testing.assert_allclose(wk, s2(t - k), atol=2e-15, err_msg=f"node c{k:+d}")
    np.testing.assert_allclose(sum(w.values()), 1.0, atol=4e-15)  # partition of unity: charge is conserved by construction


def test_cumulative_weights_of_the_current_deposit():
    """C(node) = sum of the weights up to that node; J_x(node) = -(q/dt) sum_particles [C(t_new) - C(t_old)]."""
    t = T
    P, N = np.maximum(t - 0.5, 0) ** 2, np.maximum(-t - 0.5, 0) ** 2
    C = {-2: N / 2, -1: ((t - 0.5) ** 2 - P) / 2 - N, 0: (1.5 - t) ** 2 / 2 - (t - 0.5) ** 2 + P + N / 2, 1: 1 - P / 2}
    run = np.zeros_like(t)
    for k in (-2, -1, 0, 1):
        run = run + s2(t - k)
        np.testing.assert_allclose(C[k], run, atol=4e-15, err_msg=f"cumulative weight at c{k:+d}")
    # for the particle's old position |t_old| <= 1/2 both truncated powers vanish, as the header says
    old = np.linspace(-0.5, 0.5, 101)
    assert not np.maximum(old - 0.5, 0).any() and not np.maximum(-old - 0.5, 0).any()


def test_face_weights_from_moments():
    """rho(x_n) on the faces c-3 .. c+2 (face k sits at node k + 1/2); knots at -1, 0, 1."""
    ts = T
    Nm, Z, Pp = np.maximum(-ts - 1, 0) ** 2, np.maximum(ts, 0) ** 2, np.maximum(ts - 1, 0) ** 2
    W = {-3: Nm / 2, -2: (ts ** 2 - 3 * Nm - Z) / 2, -1: (1 - 2 * ts - 2 * ts ** 2 + 3 * Nm + 3 * Z - Pp) / 2,
         0: ((ts + 1) ** 2 - Nm - 3 * Z + 3 * Pp) / 2, 1: (Z - 3 * Pp) / 2, 2: Pp / 2}
    for k, wk in W.items():
        np.testing.assert_allclose(wk, s2(ts - (k + 0.5)), atol=4e-15, err_msg=f"face c{k:+d}")
    np.testing.assert_allclose(sum(W.values()), 1.0, atol=6e-15)


@pytest.mark.parametrize("seed", [0, 1])
def test_moment_sums_reproduce_a_direct_deposit(seed):
    """A bin of particles deposited node by node == the node values formed from the five moment sums (what one work item does)."""
    rng = np.random.default_rng(seed)
    t = rng.uniform(-1.45, 1.45, 5000)
    a = rng.normal(size=t.size)  # per-particle amplitude (q, q v_y, ...)
    direct = {k: np.sum(a * s2(t - k)) for k in range(-2, 3)}
    P, N = np.maximum(t - 0.5, 0) ** 2, np.maximum(-t - 0.5, 0) ** 2
    S1, St, Stt, SP, SN = a.sum(), (a * t).sum(), (a * t * t).sum(), (a * P).sum(), (a * N).sum()
    S_tm = Stt - St + 0.25 * S1   # sum a (t - 1/2)^2
    S_tp = Stt + St + 0.25 * S1   # sum a (t + 1/2)^2
    from_moments = {-2: SN / 2, -1: (S_tm - SP - 3 * SN) / 2, 0: 0.75 * S1 - Stt + 1.5 * (SN + SP), 1: (S_tp - SN - 3 * SP) / 2, 2: SP / 2}
    for k in direct:
        assert abs(from_moments[k] - direct[k]) < 1e-11 * np.abs(a).sum()
