"""The compiled oracle (oracle/c/jic_oracle.c through oracle/c_port.py) against the NumPy oracles and the golden vectors.

It restates oracle/closed_form.py per particle with the same expressions in the same order, so single-threaded runs agree with it
to a few ulp of the grid sums; with OpenMP only the order of the grid reduction changes."""
import glob
import os

import numpy as np
import pytest

from oracle import c_port as CP
from oracle import closed_form as C
from plasma import cfl_dt, two_species

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = ("electric_field", "magnetic_field", "current_density", "charge_density", "positions", "velocities")
GOLDEN = [f for f in sorted(glob.glob(os.path.join(HERE, "golden", "*.npz"))) if "crank_nicolson" not in f]


def relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def test_library_builds_and_loads():
    assert os.path.exists(CP.build())
    assert CP.load().jo_abi_version() == 2


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(f)[:-4] for f in GOLDEN])
@pytest.mark.parametrize("threads", [1, 3])
def test_golden_vectors(path, threads):
    """Every explicit-stepper golden case, the reference-source family (refsrc_*) included."""
    g = dict(np.load(path))
    pbl, pbr, fbl, fbr = (int(b) for b in g["bcs"])
    solver = dict(filter_passes=int(g["filter_passes"]), filter_alpha=float(g["filter_alpha"]), filter_strides=tuple(int(s) for s in g["filter_strides"]),
                  relativistic=bool(g["relativistic"]), field_solver=int(g["field_solver"]) if "field_solver" in g else 0)
    out = CP.run(g["x0"], g["v0"], g["q"], g["m"], g["qm"], length=float(g["length"]), G=int(g["G"]), dt=float(g["dt"]), total_steps=int(g["T"]),
                 pbl=pbl, pbr=pbr, fbl=fbl, fbr=fbr, solver=solver, ext_E=g["ext_E"], ext_B=g["ext_B"],
                 box_yz=tuple(float(b) for b in g["box_yz"]) if "box_yz" in g else None, threads=threads)
    for k in KEYS:
        assert relerr(out[k], g[k]) < 1e-10, k
    assert relerr(out["fields"][0], g["E0"]) < 1e-10
    assert relerr(out["initial_velocities"], g["initial_velocities"]) < 1e-15


@pytest.mark.parametrize("bcs", [(0, 0, 0, 0), (1, 1, 1, 1), (2, 2, 2, 2), (1, 2, 2, 1), (2, 0, 0, 1)])
@pytest.mark.parametrize("relativistic", [False, True])
def test_matches_the_closed_form_oracle(bcs, relativistic):
    G, length, T = 24, 0.01, 25
    p = two_species(3000, 2500, length=length, G=G, seed=21 + sum(bcs), vth_e=0.12, vth_yz=0.05, drift=3e7, plus_minus=True, gpdl=0.5)
    dt = cfl_dt(length, G, 0.95)  # (the light wave is unstable above CFL 1 and would amplify the rounding of the reduction order)
    pbl, pbr, fbl, fbr = bcs
    rng = np.random.default_rng(5)
    ext_E = (1e3 * rng.standard_normal((G, 3))).astype(np.float32)
    ext_B = (1e-3 * rng.standard_normal((G, 3))).astype(np.float32)
    solver = dict(filter_passes=3, filter_alpha=0.4, filter_strides=(1, 2), relativistic=relativistic)
    kw = dict(length=length, G=G, dt=dt, total_steps=T, pbl=pbl, pbr=pbr, fbl=fbl, fbr=fbr, solver=solver, ext_E=ext_E, ext_B=ext_B, box_yz=(0.004, 0.02))
    ref = C.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], **kw)
    one = CP.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], threads=1, **kw)
    many = CP.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], threads=4, **kw)
    for k in KEYS:
        assert relerr(one[k], ref[k]) < 1e-11, k
        assert relerr(many[k], ref[k]) < 1e-10, k
    np.testing.assert_array_equal(one["initial_velocities"], ref["initial_velocities"])
    st = ref["state"]
    assert relerr(one["x_half"], st.x_half) < 1e-12 and relerr(one["v_final"], st.v) < 1e-12


@pytest.mark.parametrize("field_solver", [1, 2, 3])
def test_field_solver_branch(field_solver):
    G, length, T = 20, 0.01, 12
    p = two_species(1500, 1200, length=length, G=G, seed=31, vth_e=0.06, vth_yz=0.02, drift=5e7, plus_minus=True, gpdl=0.6)
    dt = cfl_dt(length, G, 0.9)
    kw = dict(length=length, G=G, dt=dt, total_steps=T, solver=dict(field_solver=field_solver, filter_passes=2))
    ref = C.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], **kw)
    got = CP.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], threads=2, **kw)
    for k in KEYS:
        assert relerr(got[k], ref[k]) < 1e-10, k


def test_larger_run_against_numpy_and_charge_conservation():
    """4e5 particles, G=1024 against the NumPy closed form; 2e6 particles: total charge on the grid every step."""
    G, length, T = 1024, 0.05, 5
    dt = cfl_dt(length, G, 1.0)
    p = two_species(200_000, 200_000, length=length, G=G, seed=40, vth_e=0.05, vth_yz=0.01, drift=6e7, plus_minus=True)
    kw = dict(length=length, G=G, dt=dt, total_steps=T, keep_particles=False)
    ref = C.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], **kw)
    got = CP.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], **kw)
    for k in KEYS[:4]:
        assert relerr(got[k], ref[k]) < 1e-9, k
    p = two_species(1_000_000, 1_000_000, length=length, G=G, seed=41, vth_e=0.05, vth_yz=0.01, drift=6e7, plus_minus=True)
    out = CP.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], solver=dict(filter_passes=0), **kw)
    total = out["charge_density"].sum(axis=1) * (length / G)
    assert np.abs(total - p["q"].sum()).max() < 1e-9 * np.abs(p["q"]).sum()


@pytest.mark.parametrize("seed", range(40))
def test_random_configurations_against_the_closed_form(seed):
    """The corners of tests/test_cuda_source_on_cpu.py::_random_case (every BC combination, grids from 3 cells, filter passes beyond the
    cap, strides larger than the grid, CFL 2.5, thin transverse boxes) for the compiled oracle as well."""
    from test_cuda_source_on_cpu import _random_case
    g = _random_case(seed)
    pbl, pbr, fbl, fbr = (int(b) for b in g["bcs"])
    kw = dict(length=g["length"], G=g["G"], dt=g["dt"], total_steps=g["T"], pbl=pbl, pbr=pbr, fbl=fbl, fbr=fbr, box_yz=tuple(g["box_yz"]),
              ext_E=g["ext_E"], ext_B=g["ext_B"],
              solver=dict(filter_passes=g["filter_passes"], filter_alpha=g["filter_alpha"], filter_strides=tuple(int(s) for s in g["filter_strides"]),
                          relativistic=bool(g["relativistic"])))
    ref = C.run(g["x0"], g["v0"], g["q"], g["m"], g["qm"], **kw)
    got = CP.run(g["x0"], g["v0"], g["q"], g["m"], g["qm"], threads=2, **kw)
    for k in KEYS:
        assert relerr(got[k], ref[k]) < 1e-7, k  # (CFL 2.5 cases amplify the rounding of the reduction order)
