"""Parity of the CUDA hot path (through the C ABI) with the oracle on identical initial particles.

Tolerance (BASELINE.json north_star): per-step E, B, J, rho, x, v within rtol 1e-5 in fp64 (1e-3 in fp32), measured
against the largest magnitude of each history (atomic summation order is not deterministic).
"""
import numpy as np
import pytest
import torch

from oracle import closed_form as C
from oracle import literal as L
from plasma import cfl_dt, two_species

pytestmark = pytest.mark.gpu

RTOL = {torch.float64: 1e-5, torch.float32: 1e-3}
FIELD_KEYS = ("electric_field", "magnetic_field", "current_density", "charge_density")


def run_gpu(p, *, length, G, dt, T, bcs=(0, 0, 0, 0), solver=None, ext_E=None, ext_B=None, dtype=torch.float64,
            engine="indexed", deposit="auto", particles=True, box_yz=None, steps_per_graph=0):
    from jaxincell_b200 import HotPath
    solver = {"filter_passes": 5, "filter_alpha": 0.5, "filter_strides": (1, 2, 4), "relativistic": False, "field_solver": 0, **(solver or {})}
    hp = HotPath(species=p["species"], dtype=dtype, field_solver=solver["field_solver"], length=length, G=G, dt=dt, pbl=bcs[0], pbr=bcs[1], fbl=bcs[2], fbr=bcs[3],
                 filter_passes=solver["filter_passes"], filter_alpha=solver["filter_alpha"], filter_strides=solver["filter_strides"],
                 relativistic=solver["relativistic"], engine=engine, deposit=deposit, track_yz=particles,
                 length_y=(box_yz or (0, 0))[0], length_z=(box_yz or (0, 0))[1], steps_per_graph=steps_per_graph)
    hp.set_external_fields(ext_E, ext_B)
    hp.initialize(p["x0"], p["v0"])
    out = hp.run(T, particles=particles and engine == "indexed")
    res = {k: v.cpu().numpy().astype(np.float64) for k, v in out.items()}
    E0, B0, vi = hp.initial(velocities=engine == "indexed")
    res["fields"] = (E0.cpu().numpy(), B0.cpu().numpy())
    if vi is not None:
        res["initial_velocities"] = vi.cpu().numpy()
    res["launches"] = hp.launch_count()
    res["hp"] = hp
    return res


ROW_FLOOR = 1e-3  # a step whose field is below this fraction of the run's maximum is measured against the floor (pure noise otherwise)


def assert_parity(got, ref, keys, rtol):
    """Two norms.  Global: max |cuda - oracle| over the whole history / max |oracle| over the whole history.
    Per step (row-wise): the same ratio step by step, the denominator floored at ROW_FLOOR of the global maximum -- so that a late blow-up of a
    small component cannot hide under an early large one.  Both must be below the tolerance."""
    for k in keys:
        scale = np.abs(ref[k]).max()
        if scale == 0:
            assert np.abs(got[k]).max() == 0, k
            continue
        err = np.abs(got[k] - ref[k]).max() / scale
        assert err < rtol, f"{k}: max rel err {err:.3e} >= {rtol}"
        T = ref[k].shape[0]
        num = np.abs(got[k] - ref[k]).reshape(T, -1).max(axis=1)
        den = np.maximum(np.abs(ref[k]).reshape(T, -1).max(axis=1), ROW_FLOOR * scale)
        row = (num / den).max()
        assert row < rtol, f"{k}: per-step max rel err {row:.3e} >= {rtol} (step {int((num / den).argmax())})"


@pytest.mark.parametrize("deposit", ["global", "shared"])
@pytest.mark.parametrize("bcs", [(0, 0, 0, 0), (1, 1, 1, 1), (2, 2, 2, 2), (1, 2, 1, 2), (2, 0, 2, 0), (0, 1, 0, 1)])
def test_small_plasma_all_boundaries(bcs, deposit):
    G, length, T = 16, 0.01, 25
    p = two_species(300, 300, length=length, G=G, seed=21, vth_e=0.3, vth_yz=0.2, drift=5e7, plus_minus=True, gpdl=0.5)
    dt = cfl_dt(length, G, 0.9)
    rng = np.random.default_rng(5)
    extE = 1e3 * rng.normal(size=(G, 3)); extB = 1e-3 * rng.normal(size=(G, 3))
    solver = dict(filter_passes=3, filter_alpha=0.4, filter_strides=(1, 2))
    ref = C.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T, pbl=bcs[0], pbr=bcs[1],
                fbl=bcs[2], fbr=bcs[3], solver=solver, ext_E=extE, ext_B=extB)
    got = run_gpu(p, length=length, G=G, dt=dt, T=T, bcs=bcs, solver=solver, ext_E=extE, ext_B=extB, deposit=deposit)
    assert_parity(got, ref, FIELD_KEYS + ("positions", "velocities"), 1e-5)
    np.testing.assert_allclose(got["initial_velocities"], ref["initial_velocities"], rtol=1e-14)
    np.testing.assert_allclose(got["fields"][0], ref["fields"][0], rtol=1e-10, atol=1e-12 * np.abs(ref["fields"][0]).max())
    if 2 in bcs[:2]:
        x, v, alive = got["hp"].particles()
        assert int((alive == 0).sum()) == int((ref["state"].q == 0).sum()) > 0


def test_literal_oracle_direct():
    """Straight against the expression-level restatement (both current deposits per step, O(N*G))."""
    G, length, T = 12, 0.01, 8
    p = two_species(60, 60, length=length, G=G, seed=3, vth_e=0.2, vth_yz=0.1, drift=3e7, plus_minus=True, gpdl=0.5)
    dt = cfl_dt(length, G, 0.8)
    ref = L.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T)
    got = run_gpu(p, length=length, G=G, dt=dt, T=T)
    assert_parity(got, ref, FIELD_KEYS + ("positions", "velocities"), 1e-5)


def test_example_input_toml_regime():
    """examples/input.toml: G=70, L=0.01, CFL 4.5, 3500+3500, drift +-6e7, v_y=v_z=0 (window truncation active)."""
    G, length, T = 70, 0.01, 60
    p = two_species(3500, 3500, length=length, G=G, seed=1701, vth_e=0.05, drift=6e7, plus_minus=True, gpdl=0.50265482457,
                    amp=5e-7, k=1.0, random_x=False)
    dt = cfl_dt(length, G, 4.5)
    ref = C.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T)
    got = run_gpu(p, length=length, G=G, dt=dt, T=T)
    assert_parity(got, ref, FIELD_KEYS + ("positions", "velocities"), 1e-5)


def test_landau_regime_with_energy():
    """examples/Landau_damping.py scaled down: G=32, L=1, CFL 1, vth 0.35c, heavy ions; energy history matches."""
    G, length, T = 32, 1.0, 80
    p = two_species(4000, 4000, length=length, G=G, seed=8, vth_e=0.35, gpdl=0.4, amp=0.025, k=1.02, random_x=False, ion_mass=1e9,
                    ion_vth_scale=np.sqrt(1e-9))
    dt = cfl_dt(length, G, 1.0)
    ref = C.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T)
    got = run_gpu(p, length=length, G=G, dt=dt, T=T)
    assert_parity(got, ref, FIELD_KEYS + ("positions", "velocities"), 1e-5)
    eo = C.energies(ref, p["m"], ref["dx"]); eg = C.energies(got, p["m"], ref["dx"])
    np.testing.assert_allclose(eg["total_energy"], eo["total_energy"], rtol=1e-8)
    ke = float(got["hp"].kinetic_energy().cpu()[0])
    np.testing.assert_allclose(ke, eo["kinetic_energy"][-1], rtol=1e-10)


def test_weibel_regime_magnetic_growth():
    """examples/Weibel_instability.py scaled: anisotropic vth_z >> vth_x, random x, full v x B rotation, B_y grows."""
    G, length, T = 48, 0.3, 120
    p = two_species(6000, 6000, length=length, G=G, seed=12, vth_e=0.01, vth_yz=0.10, gpdl=1.1)
    p["v0"][:6000, 1] *= 0.0  # vth_y = 0 for electrons as in the example
    dt = cfl_dt(length, G, 1.0)
    ref = C.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T)
    got = run_gpu(p, length=length, G=G, dt=dt, T=T)
    assert_parity(got, ref, FIELD_KEYS + ("positions", "velocities"), 1e-5)
    assert np.abs(got["magnetic_field"][-1]).max() > 0


def test_relativistic_pusher():
    G, length, T = 24, 0.02, 30
    p = two_species(500, 500, length=length, G=G, seed=31, vth_e=0.5, vth_yz=0.3, gpdl=0.5)
    speed = np.linalg.norm(p["v0"], axis=1, keepdims=True)  # component-wise 0.99c clipping does not bound |v|; keep gamma real
    p["v0"] = np.where(speed > 0.95 * L.speed_of_light, p["v0"] * (0.95 * L.speed_of_light / speed), p["v0"])
    dt = cfl_dt(length, G, 0.9)
    solver = dict(relativistic=True)
    ref = C.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T, solver=solver)
    got = run_gpu(p, length=length, G=G, dt=dt, T=T, solver=solver)
    assert_parity(got, ref, FIELD_KEYS + ("positions", "velocities"), 1e-5)


@pytest.mark.parametrize("G", [4, 5, 7])
def test_tiny_grids(G):
    """Reference fixtures use G=4 (tests/helpers.py:36): the J_x window degenerates to the whole grid."""
    length, T = 0.01, 12
    p = two_species(50, 50, length=length, G=G, seed=G, vth_e=0.2, vth_yz=0.1, gpdl=0.5)
    dt = cfl_dt(length, G, 0.7)
    solver = dict(filter_passes=0)
    ref = C.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T, solver=solver)
    got = run_gpu(p, length=length, G=G, dt=dt, T=T, solver=solver)
    assert_parity(got, ref, FIELD_KEYS + ("positions", "velocities"), 1e-5)


def test_fp32_mode():
    G, length, T = 32, 0.01, 30
    p = two_species(2000, 2000, length=length, G=G, seed=5, vth_e=0.1, vth_yz=0.05, drift=4e7, plus_minus=True, gpdl=0.5)
    dt = cfl_dt(length, G, 0.9)
    ref = C.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T)
    got = run_gpu(p, length=length, G=G, dt=dt, T=T, dtype=torch.float32)
    assert_parity(got, ref, FIELD_KEYS + ("positions", "velocities"), 1e-3)


def test_history_rows_and_graph_chunks():
    """Two jic_run calls continue the same simulation; odd step counts exercise the 16-step and 1-step graphs."""
    G, length = 20, 0.01
    p = two_species(400, 400, length=length, G=G, seed=77, vth_e=0.1, drift=3e7, plus_minus=True, gpdl=0.5)
    dt = cfl_dt(length, G, 0.9)
    ref = C.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=37)
    a = run_gpu(p, length=length, G=G, dt=dt, T=19, particles=False)
    b = a["hp"].run(18)
    both = {k: np.concatenate([a[k], b[k].cpu().numpy()]) for k in FIELD_KEYS}
    assert_parity(both, ref, FIELD_KEYS, 1e-5)
    assert a["hp"].launch_count() >= 2 * 37


@pytest.mark.parametrize("engine", ["indexed", "binned"])
@pytest.mark.parametrize("chunk", [None, "37", "256"])
def test_initialize_host_chunked_upload(engine, chunk, monkeypatch):
    """jic_initialize_host: host buffers uploaded in chunks (alternating staging buffers) while the start-up kernels run; same
    result as jic_initialize on device tensors, whatever the chunk size."""
    from jaxincell_b200 import HotPath
    if chunk:
        monkeypatch.setenv("JIC_HOST_CHUNK", chunk)
    G, length, T = 16, 0.01, 10
    p = two_species(300, 211, length=length, G=G, seed=9, vth_e=0.1, vth_yz=0.05, gpdl=0.5)
    dt = cfl_dt(length, G, 0.9)
    ref = C.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T, pbl=1, pbr=2, fbl=1, fbr=2)
    hp = HotPath(species=p["species"], length=length, G=G, dt=dt, engine=engine, pbl=1, pbr=2, fbl=1, fbr=2, track_yz=engine == "indexed")
    hp.set_external_fields(None, None)
    hx, hv = torch.from_numpy(p["x0"]).pin_memory(), torch.from_numpy(p["v0"]).pin_memory()
    for _ in range(2):  # twice: the staging buffers and events are reused
        hp.initialize_host(hx, hv)
        out = hp.run(T, particles=engine == "indexed")
    got = {k: v.cpu().numpy() for k, v in out.items()}
    assert_parity(got, ref, FIELD_KEYS + (("positions", "velocities") if engine == "indexed" else ()), 1e-5)
    hp.close()


def test_host_buffer_entry_point():
    """jic_simulate_host: NumPy in, NumPy out, all copies inside (the e2e boundary)."""
    from jaxincell_b200 import simulate_host
    G, length, T = 16, 0.01, 10
    p = two_species(200, 200, length=length, G=G, seed=9, vth_e=0.1, vth_yz=0.05, gpdl=0.5)
    dt = cfl_dt(length, G, 0.9)
    ref = C.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T)
    got = simulate_host(species=p["species"], x0=p["x0"], v0=p["v0"], n_steps=T, length=length, G=G, dt=dt, particles=True, initial=True)
    assert_parity(got, ref, FIELD_KEYS + ("positions", "velocities"), 1e-5)
    np.testing.assert_allclose(got["initial_velocities"], ref["initial_velocities"], rtol=1e-14)


# ------------------------------------------------------------------------------------------------- BINNED engine
def _sorted_particles(x, v, alive):
    keep = alive.astype(bool)
    x, v = x[keep], v[keep]
    order = np.lexsort((v[:, 2], v[:, 1], v[:, 0], x[:, 0]))
    return x[order, 0], v[order]


@pytest.mark.parametrize("scatter", ["one_pass", "records"])
@pytest.mark.parametrize("bcs", [(0, 0, 0, 0), (1, 1, 1, 1), (2, 2, 2, 2), (1, 2, 1, 2), (2, 0, 2, 0)])
def test_binned_engine_all_boundaries(bcs, scatter, monkeypatch):
    """Cell-binned store (fast path + general path in wall cells + re-binning) against the oracle, fields AND particles.
    scatter: the start-up's way into the bins -- the one-pass kernel (default below 2^20 particles) or the two-pass one through
    whole-sector records (default above; forced here)."""
    monkeypatch.setenv("JIC_SCATTER_RECORDS", "2" if scatter == "records" else "0")
    G, length, T = 24, 0.01, 30
    p = two_species(3000, 3000, length=length, G=G, seed=41, vth_e=0.3, vth_yz=0.2, drift=5e7, plus_minus=True, gpdl=0.5)
    dt = cfl_dt(length, G, 0.9)
    rng = np.random.default_rng(6)
    extE = 1e3 * rng.normal(size=(G, 3)); extB = 1e-3 * rng.normal(size=(G, 3))
    ref = C.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T, pbl=bcs[0], pbr=bcs[1],
                fbl=bcs[2], fbr=bcs[3], ext_E=extE, ext_B=extB)
    got = run_gpu(p, length=length, G=G, dt=dt, T=T, bcs=bcs, ext_E=extE, ext_B=extB, engine="binned", particles=False)
    assert_parity(got, ref, FIELD_KEYS, 1e-5)
    x, v, alive = (t.cpu().numpy() for t in got["hp"].particles())
    st = ref["state"]
    gx, gv = _sorted_particles(x, v, alive)
    rx, rv = _sorted_particles(st.x_half, st.v, (st.q != 0).astype(np.uint8))
    assert len(gx) == len(rx)
    np.testing.assert_allclose(gx, rx, rtol=0, atol=1e-9 * length)
    np.testing.assert_allclose(gv, rv, rtol=1e-7, atol=1e-7 * np.abs(rv).max())
    ke_ref = 0.5 * np.sum(st.m * np.sum(st.v ** 2, axis=1) * (st.q != 0))
    np.testing.assert_allclose(float(got["hp"].kinetic_energy().cpu()[0]), ke_ref, rtol=1e-9)


def test_binned_engine_large_cfl_multi_cell_jumps():
    """CFL 4.5 (examples/input.toml): many particles jump >1 cell -> general path and window truncation inside the binned push."""
    G, length, T = 70, 0.01, 40
    p = two_species(7000, 7000, length=length, G=G, seed=1701, vth_e=0.05, drift=6e7, plus_minus=True, gpdl=0.50265482457,
                    amp=5e-7, k=1.0, random_x=False)
    dt = cfl_dt(length, G, 4.5)
    ref = C.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T)
    got = run_gpu(p, length=length, G=G, dt=dt, T=T, engine="binned", particles=False)
    assert_parity(got, ref, FIELD_KEYS, 1e-5)


def test_binned_engine_relativistic_and_fp32(monkeypatch):
    monkeypatch.setenv("JIC_SCATTER_RECORDS", "2")  # (the two-pass start-up scatter with 16-byte records in the fp32 leg)
    G, length, T = 32, 0.02, 25
    p = two_species(4000, 4000, length=length, G=G, seed=31, vth_e=0.4, vth_yz=0.2, gpdl=0.5)
    speed = np.linalg.norm(p["v0"], axis=1, keepdims=True)
    p["v0"] = np.where(speed > 0.95 * L.speed_of_light, p["v0"] * (0.95 * L.speed_of_light / speed), p["v0"])
    dt = cfl_dt(length, G, 0.9)
    solver = dict(relativistic=True)
    ref = C.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T, solver=solver)
    got = run_gpu(p, length=length, G=G, dt=dt, T=T, solver=solver, engine="binned", particles=False)
    assert_parity(got, ref, FIELD_KEYS, 1e-5)
    ref = C.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T)
    got = run_gpu(p, length=length, G=G, dt=dt, T=T, engine="binned", particles=False, dtype=torch.float32)
    assert_parity(got, ref, FIELD_KEYS, 1e-3)


def test_binned_engine_nonuniform_density_overflow_path():
    """All particles start in a quarter of the box and stream out: bins outgrow their head-room, the overflow list carries them."""
    G, length, T = 64, 0.01, 40
    p = two_species(20000, 20000, length=length, G=G, seed=3, vth_e=0.2, vth_yz=0.05, gpdl=0.3)
    p["x0"][:, 0] = p["x0"][:, 0] / 4.0
    dt = cfl_dt(length, G, 0.9)
    ref = C.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T, keep_particles=False)
    got = run_gpu(p, length=length, G=G, dt=dt, T=T, engine="binned", particles=False)
    assert_parity(got, ref, FIELD_KEYS, 1e-5)
    x, v, alive = (t.cpu().numpy() for t in got["hp"].particles())
    assert int(alive.sum()) == 40000


def test_binned_matches_indexed_at_scale():
    """2e6 particles, G=512: the two engines must agree with each other far beyond what the oracle can check quickly."""
    G, length, T = 512, 0.05, 20
    p = two_species(1_000_000, 1_000_000, length=length, G=G, seed=5, vth_e=0.05, vth_yz=0.01, drift=6e7, plus_minus=True)
    dt = cfl_dt(length, G, 1.0)
    a = run_gpu(p, length=length, G=G, dt=dt, T=T, engine="indexed", particles=False)
    b = run_gpu(p, length=length, G=G, dt=dt, T=T, engine="binned", particles=False)
    assert_parity(b, a, FIELD_KEYS, 1e-7)


@pytest.mark.parametrize("engine", ["indexed", "binned"])
@pytest.mark.parametrize("G,bcs,solver", [
    (300, (0, 0, 0, 0), {}),                                                   # 4 slices of 75 nodes, halo 37
    (333, (0, 0, 0, 0), dict(filter_passes=2, filter_strides=(1, 3))),         # uneven last slice
    (300, (1, 1, 1, 1), {}),
    (300, (2, 2, 2, 2), {}),
    (310, (1, 2, 1, 2), dict(filter_passes=3, filter_alpha=0.4, filter_strides=(2, 5))),
    (300, (2, 0, 2, 0), dict(filter_passes=0)),                                # no filter: halo 2
    (4096, (0, 0, 0, 0), {}),                                                  # the bench grid: 16 slices of 256
])
def test_multi_cta_field_kernel(G, bcs, solver, engine):
    """Grids large enough for k_fields_mc (slices + recomputed halos, ping-ponged state) against the oracle, every BC."""
    length, T = 0.05, 14
    p = two_species(12000, 9000, length=length, G=G, seed=77, vth_e=0.05, vth_yz=0.03, drift=5e7, plus_minus=True, gpdl=2.0)
    dt = cfl_dt(length, G, 0.9)
    rng = np.random.default_rng(8)
    extE = 1e2 * rng.normal(size=(G, 3)); extB = 1e-4 * rng.normal(size=(G, 3))
    ref = C.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T, pbl=bcs[0], pbr=bcs[1],
                fbl=bcs[2], fbr=bcs[3], solver=solver, ext_E=extE, ext_B=extB, keep_particles=False)
    # odd number of steps per graph and a second call: exercises both parities of the ping-pong buffers
    from jaxincell_b200 import HotPath
    s = {"filter_passes": 5, "filter_alpha": 0.5, "filter_strides": (1, 2, 4), **solver}
    hp = HotPath(species=p["species"], length=length, G=G, dt=dt, pbl=bcs[0], pbr=bcs[1], fbl=bcs[2], fbr=bcs[3],
                 filter_passes=s["filter_passes"], filter_alpha=s["filter_alpha"], filter_strides=s["filter_strides"],
                 engine=engine, steps_per_graph=3)
    hp.set_external_fields(extE, extB)
    hp.initialize(p["x0"], p["v0"])
    a = hp.run(5)
    b = hp.run(T - 5)
    got = {k: torch.cat([a[k], b[k]]).cpu().numpy() for k in FIELD_KEYS}
    assert_parity(got, ref, FIELD_KEYS, 1e-5)
    E, B, J, rho = (t.cpu().numpy() for t in hp.fields())
    np.testing.assert_array_equal(E, got["electric_field"][-1])
    np.testing.assert_array_equal(rho, got["charge_density"][-1])
    hp.close()


# ------------------------------------------------------------------- per-step electrostatic correction (field_solver != 0)
@pytest.mark.parametrize("engine", ["indexed", "binned"])
@pytest.mark.parametrize("field_solver", [1, 2, 3])
@pytest.mark.parametrize("bcs", [(0, 0, 0, 0), (1, 1, 1, 1), (2, 2, 2, 2)])
def test_field_solver_correction(bcs, field_solver, engine):
    """jaxincell/_algorithms.py:69-78 (tests/test_algorithms.py:543-614 of the reference): every step E_x is replaced by the
    Gauss-FFT / Gauss-Cartesian / Poisson-FFT solve of rho(x_n) deposited on the faces with the post-BC charges."""
    G, length, T = 24, 0.01, 20
    p = two_species(2500, 2500, length=length, G=G, seed=13, vth_e=0.3, vth_yz=0.2, drift=5e7, plus_minus=True, gpdl=0.5)
    dt = cfl_dt(length, G, 0.9)
    solver = dict(field_solver=field_solver, filter_passes=3, filter_alpha=0.4, filter_strides=(1, 2))
    ref = C.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T, pbl=bcs[0], pbr=bcs[1],
                fbl=bcs[2], fbr=bcs[3], solver=solver, keep_particles=engine == "indexed")
    got = run_gpu(p, length=length, G=G, dt=dt, T=T, bcs=bcs, solver=solver, engine=engine, particles=engine == "indexed")
    assert_parity(got, ref, FIELD_KEYS + (("positions", "velocities") if engine == "indexed" else ()), 1e-5)


def test_field_solver_literal_oracle_direct():
    G, length, T = 12, 0.01, 8
    p = two_species(60, 60, length=length, G=G, seed=3, vth_e=0.2, vth_yz=0.1, drift=3e7, plus_minus=True, gpdl=0.5)
    dt = cfl_dt(length, G, 0.8)
    ref = L.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T, solver=dict(field_solver=1))
    got = run_gpu(p, length=length, G=G, dt=dt, T=T, solver=dict(field_solver=1))
    assert_parity(got, ref, FIELD_KEYS + ("positions", "velocities"), 1e-5)


@pytest.mark.parametrize("dtype,rtol", [(torch.float64, 1e-5), (torch.float32, 1e-3)])
def test_field_solver_large_grid_binned_fast_path(dtype, rtol):
    """G = 1024: multi-CTA field kernel + k_gauss over many CTAs + the moment form of the face deposit (bins away from the
    domain ends), engines against each other and against the oracle."""
    G, length, T = 1024, 0.05, 10
    p = two_species(150_000, 150_000, length=length, G=G, seed=5, vth_e=0.05, vth_yz=0.01, drift=6e7, plus_minus=True)
    dt = cfl_dt(length, G, 1.0)
    solver = dict(field_solver=1)
    ref = C.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T, solver=solver, keep_particles=False)
    b = run_gpu(p, length=length, G=G, dt=dt, T=T, engine="binned", particles=False, solver=solver, dtype=dtype, steps_per_graph=3)
    assert_parity(b, ref, FIELD_KEYS, rtol)
    if dtype == torch.float64:
        a = run_gpu(p, length=length, G=G, dt=dt, T=T, engine="indexed", particles=False, solver=solver)
        assert_parity(b, a, FIELD_KEYS, 1e-7)


# ------------------------------------------------------------------------------------------------- edge cases
@pytest.mark.parametrize("engine", ["indexed", "binned"])
def test_ragged_and_tiny_species(engine):
    """Species of very different sizes, one of them a single particle, on the smallest grid the library accepts (G = 3), and an
    EMPTY species block in the middle of the table (count 0: nothing to push, must not disturb the others)."""
    G, length, T = 3, 0.01, 10
    p = two_species(1, 37, length=length, G=G, seed=4, vth_e=0.2, vth_yz=0.1, gpdl=0.5)
    dt = cfl_dt(length, G, 0.7)
    ref = C.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T, solver=dict(filter_passes=1))
    sp = [p["species"][0], dict(count=0, q=1.0, m=1.0, qm=1.0), p["species"][1]]
    got = run_gpu(dict(p, species=sp), length=length, G=G, dt=dt, T=T, engine=engine, particles=engine == "indexed", solver=dict(filter_passes=1))
    assert_parity(got, ref, FIELD_KEYS + (("positions", "velocities") if engine == "indexed" else ()), 1e-5)


@pytest.mark.parametrize("engine", ["indexed", "binned"])
def test_no_particles_at_all(engine):
    """N = 0: the fields stay zero, nothing crashes."""
    from jaxincell_b200 import HotPath
    hp = HotPath(species=[dict(count=0, q=1.0, m=1.0, qm=1.0)], length=0.01, G=16, dt=1e-12, engine=engine)
    hp.set_external_fields(None, None)
    hp.initialize(np.zeros((0, 3)), np.zeros((0, 3)))
    out = hp.run(5)
    torch.cuda.synchronize()
    assert all(float(v.abs().max()) == 0.0 for v in out.values())
    hp.close()


def test_all_particles_in_one_cell_binned():
    """Maximum collision case of the binned store: every particle of both species starts in the same cell (one bin holds all)."""
    G, length, T = 32, 0.01, 15
    p = two_species(5000, 5000, length=length, G=G, seed=6, vth_e=0.1, vth_yz=0.05, gpdl=0.3)
    dx = length / G
    p["x0"][:, 0] = 0.25 * dx + 0.5 * dx * (p["x0"][:, 0] / length)   # all inside cell G/2
    dt = cfl_dt(length, G, 0.9)
    ref = C.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T, keep_particles=False)
    got = run_gpu(p, length=length, G=G, dt=dt, T=T, engine="binned", particles=False)
    assert_parity(got, ref, FIELD_KEYS, 1e-5)
    x, v, alive = (t.cpu().numpy() for t in got["hp"].particles())
    assert int(alive.sum()) == 10000


def test_device_memory_pool_keeps_freed_buffers_until_trimmed():
    """The large buffers of a context come from a library-owned memory pool that keeps them after jic_destroy (the next context of the
    process starts without cudaMalloc); jic_trim_memory hands them back."""
    from jaxincell_b200 import HotPath, trim_memory
    trim_memory()
    G, length = 256, 0.01
    p = two_species(1_000_000, 1_000_000, length=length, G=G, seed=3, vth_e=0.05, vth_yz=0.01, drift=3e7, plus_minus=True, gpdl=0.5)
    dt = cfl_dt(length, G, 0.9)
    x0, v0 = torch.as_tensor(p["x0"], device="cuda"), torch.as_tensor(p["v0"], device="cuda")
    torch.cuda.synchronize()
    free_before = torch.cuda.mem_get_info()[0]
    fields = []
    for _ in range(2):  # the second context runs on recycled memory: same result
        hp = HotPath(species=p["species"], length=length, G=G, dt=dt, engine="binned")
        hp.set_external_fields(None, None)
        hp.initialize(x0, v0)
        fields.append(hp.run(3)["electric_field"].cpu().numpy())
        hp.check_status()
        hp.close()
    torch.cuda.synchronize()
    kept = free_before - torch.cuda.mem_get_info()[0]
    assert kept > 100 << 20, kept          # 2 x 2.25 x 2e6 slots x 32 B of particle store stay with the pool
    trim_memory()
    assert free_before - torch.cuda.mem_get_info()[0] < 32 << 20
    scale = np.abs(fields[0]).max()
    assert np.abs(fields[1] - fields[0]).max() < 1e-9 * scale
