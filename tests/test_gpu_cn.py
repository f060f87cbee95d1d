"""Implicit Crank-Nicolson stepper (jaxincell/_algorithms.py:100-241) on the GPU against the oracle's
CN_step: same Picard iteration counts, fields and particles within the north-star tolerance (1e-5 fp64, 1e-3 fp32)."""
import numpy as np
import pytest
import torch

from oracle import closed_form as C
from oracle import literal as L
from plasma import cfl_dt, two_species

pytestmark = pytest.mark.gpu
KEYS = ("electric_field", "magnetic_field", "current_density", "charge_density", "positions", "velocities")


@pytest.fixture
def sorted_push(monkeypatch):
    """Contexts created inside the test take the cell-sorted Crank-Nicolson push (csrc/jic_cn_sorted.cuh) whatever their size: the
    library reads the threshold (default 200000 particles) when the context is created."""
    monkeypatch.setenv("JIC_CN_SORTED_MIN", "0")


def run_gpu_cn(p, *, length, G, dt, T, bcs=(0, 0, 0, 0), solver=None, dtype=torch.float64, steps_per_graph=0, split=None, deposit="auto", kinetic=False):
    from jaxincell_b200 import HotPath
    s = {"max_number_of_Picard_iterations_implicit_CN": 20, "number_of_particle_substeps_implicit_CN": 2,
         "tolerance_Picard_iterations_implicit_CN": 1e-6, "filter_passes": 5, "filter_alpha": 0.5, "filter_strides": (1, 2, 4), **(solver or {})}
    hp = HotPath(species=p["species"], dtype=dtype, length=length, G=G, dt=dt, pbl=bcs[0], pbr=bcs[1], fbl=bcs[2], fbr=bcs[3],
                 filter_passes=s["filter_passes"], filter_alpha=s["filter_alpha"], filter_strides=s["filter_strides"],
                 time_evolution_algorithm=1, cn_substeps=s["number_of_particle_substeps_implicit_CN"],
                 cn_max_iterations=s["max_number_of_Picard_iterations_implicit_CN"], cn_tolerance=s["tolerance_Picard_iterations_implicit_CN"],
                 steps_per_graph=steps_per_graph, deposit=deposit)
    hp.set_external_fields(None, None)
    hp.initialize(p["x0"], p["v0"])
    iters = []
    if split:
        a = hp.run(split, particles=True, kinetic=kinetic); iters.append(hp.picard_iterations())
        b = hp.run(T - split, particles=True, kinetic=kinetic)
        out = {k: torch.cat([a[k], b[k]]) for k in a}
    else:
        out = hp.run(T, particles=True, kinetic=kinetic)
    iters.append(hp.picard_iterations())
    res = {k: v.cpu().numpy().astype(np.float64) for k, v in out.items()}
    E0, B0, vi = hp.initial()
    res.update(fields=(E0.cpu().numpy(), B0.cpu().numpy()), initial_velocities=vi.cpu().numpy(), iters=iters, hp=hp)
    return res


def assert_parity(got, ref, rtol, keys=KEYS):
    for k in keys:
        scale = np.abs(ref[k]).max()
        err = np.abs(got[k] - ref[k]).max() / max(scale, 1e-300)
        assert err < rtol, f"{k}: max rel err {err:.3e} >= {rtol}"


@pytest.mark.parametrize("deposit", ["global", "shared"])
@pytest.mark.parametrize("bcs", [(0, 0, 0, 0), (1, 1, 1, 1), (2, 2, 2, 2), (1, 2, 1, 2)])
def test_cn_matches_the_oracle_all_boundaries(bcs, deposit):
    G, length, T = 24, 0.01, 12
    p = two_species(1500, 1500, length=length, G=G, seed=13, vth_e=0.2, vth_yz=0.1, drift=4e7, plus_minus=True, gpdl=0.008)
    dt = cfl_dt(length, G, 0.3)
    solver = dict(tolerance_Picard_iterations_implicit_CN=1e-9, max_number_of_Picard_iterations_implicit_CN=12, number_of_particle_substeps_implicit_CN=3)
    ref = L.run_CN(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T, pbl=bcs[0], pbr=bcs[1], fbl=bcs[2],
                   fbr=bcs[3], solver=solver)
    got = run_gpu_cn(p, length=length, G=G, dt=dt, T=T, bcs=bcs, solver=solver, steps_per_graph=5, split=7, deposit=deposit)
    assert_parity(got, ref, 1e-5)
    np.testing.assert_allclose(got["initial_velocities"], ref["initial_velocities"], rtol=1e-14)
    np.testing.assert_allclose(got["fields"][0], ref["fields"][0], rtol=1e-10, atol=1e-12 * np.abs(ref["fields"][0]).max())
    assert got["iters"][-1][0] == ref["picard_iterations"][-1]
    assert got["iters"][-1][1] == ref["picard_iterations"].sum()
    assert got["iters"][0][1] == ref["picard_iterations"][:7].sum()


@pytest.mark.parametrize("tol,max_iter,substeps", [(1e9, 5, 1), (1e-30, 1, 1), (1e-3, 2, 3), (1e-30, 4, 2)])
def test_cn_picard_stopping_rules(tol, max_iter, substeps):
    """reference tests/test_algorithms.py:616-680: early exit on tolerance, cap on iterations, sub-step count."""
    G, length, T = 16, 0.01, 5
    p = two_species(400, 400, length=length, G=G, seed=2, vth_e=0.1, vth_yz=0.05, drift=3e7, plus_minus=True, gpdl=0.05)
    dt = cfl_dt(length, G, 0.8)
    solver = dict(tolerance_Picard_iterations_implicit_CN=tol, max_number_of_Picard_iterations_implicit_CN=max_iter,
                  number_of_particle_substeps_implicit_CN=substeps)
    ref = L.run_CN(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T, solver=solver)
    got = run_gpu_cn(p, length=length, G=G, dt=dt, T=T, solver=solver)
    assert got["iters"][-1][1] == ref["picard_iterations"].sum()
    assert_parity(got, ref, 1e-5)


@pytest.mark.parametrize("bcs", [(0, 0, 0, 0), (1, 1, 1, 1), (2, 2, 2, 2), (1, 2, 1, 2)])
def test_cn_sorted_push_matches_the_oracle_all_boundaries(bcs, sorted_push):
    """The large-run variant (counting sort by cell every step, deposit aggregated per warp in shared memory) on the same cases: the
    histories come back in input order through the permutation, the Picard counts are the oracle's, the kinetic-energy history is
    0.5 m v^2 of the velocity history."""
    G, length, T = 24, 0.01, 12
    p = two_species(1500, 1500, length=length, G=G, seed=13, vth_e=0.2, vth_yz=0.1, drift=4e7, plus_minus=True, gpdl=0.008)
    dt = cfl_dt(length, G, 0.3)
    solver = dict(tolerance_Picard_iterations_implicit_CN=1e-9, max_number_of_Picard_iterations_implicit_CN=12, number_of_particle_substeps_implicit_CN=3)
    ref = L.run_CN(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T, pbl=bcs[0], pbr=bcs[1], fbl=bcs[2],
                   fbr=bcs[3], solver=solver)
    got = run_gpu_cn(p, length=length, G=G, dt=dt, T=T, bcs=bcs, solver=solver, steps_per_graph=5, split=7, kinetic=True)
    assert got["hp"].store_stats()["cn_sorted"] == 1
    assert_parity(got, ref, 1e-5)
    assert got["iters"][-1][0] == ref["picard_iterations"][-1]
    assert got["iters"][-1][1] == ref["picard_iterations"].sum()
    assert got["iters"][0][1] == ref["picard_iterations"][:7].sum()
    v, m = got["velocities"], p["m"]
    ne = p["species"][0]["count"]
    want = np.stack([0.5 * (m[:ne] * (v[:, :ne] ** 2).sum(-1)).sum(-1), 0.5 * (m[ne:] * (v[:, ne:] ** 2).sum(-1)).sum(-1)], axis=1)
    np.testing.assert_allclose(got["kinetic_energy"], want, rtol=1e-12)
    x, vv, alive = (t.cpu().numpy() for t in got["hp"].particles())
    np.testing.assert_allclose(x, ref["positions"][-1], rtol=0, atol=1e-9 * length)
    np.testing.assert_allclose(vv, ref["velocities"][-1], rtol=1e-7, atol=1e-9 * np.abs(ref["velocities"]).max())
    ke = float(got["hp"].kinetic_energy().cpu()[0])
    np.testing.assert_allclose(ke, want[-1].sum(), rtol=1e-10)


def test_cn_sorted_push_fp32(sorted_push):
    """The float instantiation of the sorted push (float window, float2 reduction) within the fp32 tolerance on E, J and rho."""
    G, length, T = 32, 0.01, 10
    p = two_species(4000, 4000, length=length, G=G, seed=5, vth_e=0.05, vth_yz=0.02, drift=5e7, plus_minus=True, gpdl=0.03)
    dt = cfl_dt(length, G, 0.9)
    solver = dict(tolerance_Picard_iterations_implicit_CN=1e-4, max_number_of_Picard_iterations_implicit_CN=25)
    ref = L.run_CN(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T, solver=solver)
    got = run_gpu_cn(p, length=length, G=G, dt=dt, T=T, solver=solver, dtype=torch.float32)
    assert got["hp"].store_stats()["cn_sorted"] == 1
    assert_parity(got, ref, 1e-3, keys=("electric_field", "current_density", "charge_density"))


def test_cn_kinetic_energy_history_unsorted():
    """jic_outputs.kinetic_energy of the unsorted stepper: the row is the step's (the field kernel advances the row counter first)."""
    G, length, T = 16, 0.01, 6
    p = two_species(400, 300, length=length, G=G, seed=2, vth_e=0.1, vth_yz=0.05, drift=3e7, plus_minus=True, gpdl=0.05)
    got = run_gpu_cn(p, length=length, G=G, dt=cfl_dt(length, G, 0.8), T=T, kinetic=True, split=2)
    assert got["hp"].store_stats()["cn_sorted"] == 0
    v, m = got["velocities"], p["m"]
    want = np.stack([0.5 * (m[:400] * (v[:, :400] ** 2).sum(-1)).sum(-1), 0.5 * (m[400:] * (v[:, 400:] ** 2).sum(-1)).sum(-1)], axis=1)
    np.testing.assert_allclose(got["kinetic_energy"], want, rtol=1e-12)


def test_cn_sorted_push_many_cells_wide_warps(sorted_push):
    """Few particles per cell: a warp spans more cells than its shared-memory window holds, so most contributions take the direct
    path -- the result must not depend on how good the order is.  (The Picard iteration on the fields diverges beyond the light
    CFL, in the reference as well, so particles cannot be made to jump many cells per step.)"""
    G, length, T = 256, 0.05, 8
    p = two_species(700, 650, length=length, G=G, seed=21, vth_e=0.3, vth_yz=0.1, drift=1e8, plus_minus=True, gpdl=0.02)
    dt = cfl_dt(length, G, 0.9)
    solver = dict(tolerance_Picard_iterations_implicit_CN=1e-8, max_number_of_Picard_iterations_implicit_CN=15, number_of_particle_substeps_implicit_CN=2)
    ref = L.run_CN(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T, solver=solver)
    assert np.isfinite(ref["electric_field"]).all()
    got = run_gpu_cn(p, length=length, G=G, dt=dt, T=T, solver=solver)
    assert got["hp"].store_stats()["cn_sorted"] == 1
    assert_parity(got, ref, 1e-5)
    assert got["iters"][-1][1] == ref["picard_iterations"].sum()


def test_cn_conserves_energy_and_fp32_mode():
    G, length, T = 32, 0.01, 40
    p = two_species(4000, 4000, length=length, G=G, seed=5, vth_e=0.05, vth_yz=0.02, drift=5e7, plus_minus=True, gpdl=0.03)
    dt = cfl_dt(length, G, 0.9)
    solver = dict(tolerance_Picard_iterations_implicit_CN=1e-10, max_number_of_Picard_iterations_implicit_CN=25)
    ref = L.run_CN(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T, solver=solver)
    got = run_gpu_cn(p, length=length, G=G, dt=dt, T=T, solver=solver)
    assert_parity(got, ref, 1e-5)
    e = C.energies(got, p["m"], ref["dx"])["total_energy"]
    assert abs(e[-1] / e[0] - 1) < 1e-5
    ke = float(got["hp"].kinetic_energy().cpu()[0])
    np.testing.assert_allclose(ke, C.energies(ref, p["m"], ref["dx"])["kinetic_energy"][-1], rtol=1e-8)
    x, v, alive = (t.cpu().numpy() for t in got["hp"].particles())
    np.testing.assert_allclose(x, ref["positions"][-1], rtol=0, atol=1e-9 * length)
    got32 = run_gpu_cn(p, length=length, G=G, dt=dt, T=10, solver=dict(solver, tolerance_Picard_iterations_implicit_CN=1e-4), dtype=torch.float32)
    ref10 = {k: ref[k][:10] for k in KEYS}
    # (B is ~1e-8 of E/c here -- pure fp32 particle noise -- so the fp32 check is on E, J and rho)
    assert_parity(got32, ref10, 1e-3, keys=("electric_field", "current_density", "charge_density"))


def test_cn_through_the_simulation_driver():
    from jaxincell_b200 import Simulation
    G, length = 16, 0.01
    p = two_species(300, 200, length=length, G=G, seed=9, vth_e=0.05, vth_yz=0.02, drift=3e7, plus_minus=True, gpdl=0.05)
    par = {"domain_parameters": {"number_grid_points": G, "total_steps": 6, "length": length, "timestep_over_spatialstep_times_c": 0.9},
           "species_parameters": {"electrons": {"e": {"number_pseudoparticles": 300, "vth_over_c_x": 0.05, "grid_points_per_Debye_length": 0.05,
                                                      "initial_positions": p["x0"][:300], "initial_velocities": p["v0"][:300]}},
                                  "ions": {"i": {"number_pseudoparticles": 200, "grid_points_per_Debye_length": 0.05,
                                                 "initial_positions": p["x0"][300:], "initial_velocities": p["v0"][300:]}}},
           "solver_parameters": {"print_info": False, "time_evolution_algorithm": 1, "max_number_of_Picard_iterations_implicit_CN": 8,
                                 "number_of_particle_substeps_implicit_CN": 2, "tolerance_Picard_iterations_implicit_CN": 1e-8}}
    out = Simulation(par).run()
    q, m = out["charges"].reshape(-1), out["masses"].reshape(-1)
    ref = L.run_CN(p["x0"], p["v0"], q, m, out["charge_to_mass_ratios"].reshape(-1), length=length, G=G, dt=out["dt"], total_steps=6,
                   solver=par["solver_parameters"])
    assert_parity(out, ref, 1e-5)
