"""Committed golden vectors, two families with the same keys:

  tests/golden/refsrc_<case>.npz -- written by tests/golden/make_reference_golden.py: the REFERENCE'S OWN SOURCE
      (`jaxincell.Simulation(...).run()`, i.e. its start-up, `Boris_step` / `CN_step`, scan and output dict) executed in the build
      container on the NumPy stand-in for jax of tests/refshim (jax itself is not installable); the stand-in is pinned by the
      reference's own unit tests (tests/golden/REFERENCE_SOURCE_RUN.md).  These pin the *composition* of the step.
  tests/golden/<case>.npz        -- written by tests/golden/make_golden.py from oracle/literal.py (same inputs).

CPU: both oracles must reproduce both families (guards the oracles against drift, and the restatement against the reference).
GPU: both CUDA engines, through the C ABI, must reproduce them within the north-star tolerance (1e-5 fp64)."""
import glob
import os

import numpy as np
import pytest

from oracle import closed_form as C
from oracle import literal as L

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))
FIELDS = ("electric_field", "magnetic_field", "current_density", "charge_density")
RTOL_F64 = 1e-5  # BASELINE.json north_star: per-step E, B, J and particle x/v within 1e-5 relative in fp64


def _load(path):
    g = dict(np.load(path))
    g["solver"] = dict(filter_passes=int(g["filter_passes"]), filter_alpha=float(g["filter_alpha"]),
                       filter_strides=tuple(int(s) for s in g["filter_strides"]), relativistic=bool(g["relativistic"]),
                       field_solver=int(g["field_solver"]) if "field_solver" in g else 0)
    g["cn"] = int(g["time_evolution_algorithm"]) == 1 if "time_evolution_algorithm" in g else False
    g["box_yz"] = tuple(float(b) for b in g["box_yz"]) if "box_yz" in g else None
    if g["cn"]:
        g["solver"].update(max_number_of_Picard_iterations_implicit_CN=int(g["cn_max_iterations"]),
                           number_of_particle_substeps_implicit_CN=int(g["cn_substeps"]),
                           tolerance_Picard_iterations_implicit_CN=float(g["cn_tolerance"]))
    return g


def _relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def test_fixtures_exist():
    assert len(FILES) >= 20
    assert sum(os.path.basename(f).startswith("refsrc_") for f in FILES) >= 10


REFSRC = [f for f in FILES if os.path.basename(f).startswith("refsrc_")]


@pytest.mark.parametrize("path", REFSRC, ids=[os.path.basename(f)[:-4] for f in REFSRC])
def test_reference_source_vectors_match_the_oracle_made_ones(path):
    """Same inputs, one file from the reference's source, one from oracle/literal.py: equal to round-off, Picard counts exactly."""
    r, o = dict(np.load(path)), dict(np.load(path.replace("refsrc_", "")))
    for k in ("x0", "v0", "q", "m", "qm", "ext_E", "ext_B", "bcs", "filter_strides"):
        assert np.array_equal(r[k], o[k]), k
    assert float(r["dt"]) == float(o["dt"])
    for k in FIELDS + ("positions", "velocities", "E0", "initial_velocities"):
        assert _relerr(o[k], r[k]) < 1e-10, k
    if "picard_iterations" in r:
        assert r["picard_iterations"].tolist() == o["picard_iterations"].tolist()


@pytest.mark.parametrize("path", REFSRC, ids=[os.path.basename(f)[:-4] for f in REFSRC])
def test_literal_oracle_reproduces_reference_source_vectors(path):
    g = _load(path)
    pbl, pbr, fbl, fbr = (int(b) for b in g["bcs"])
    kw = dict(length=float(g["length"]), G=int(g["G"]), dt=float(g["dt"]), total_steps=int(g["T"]), pbl=pbl, pbr=pbr, fbl=fbl, fbr=fbr, solver=g["solver"], box_yz=g["box_yz"])
    if g["cn"]:
        out = L.run_CN(g["x0"], g["v0"], g["q"], g["m"], g["qm"], **kw)
        assert out["picard_iterations"].tolist() == g["picard_iterations"].tolist()
    else:
        out = L.run(g["x0"], g["v0"], g["q"], g["m"], g["qm"], ext_E=g["ext_E"], ext_B=g["ext_B"], **kw)
    for k in FIELDS + ("positions", "velocities"):
        assert _relerr(out[k], g[k]) < 1e-10, k
    assert _relerr(out["fields"][0], g["E0"]) < 1e-10  # the reference solves a dense bidiagonal system, the oracle sums: cancellation of a neutral plasma
    assert _relerr(out["initial_velocities"], g["initial_velocities"]) < 1e-15


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_closed_form_oracle_reproduces_golden(path):
    g = _load(path)
    pbl, pbr, fbl, fbr = (int(b) for b in g["bcs"])
    kw = dict(length=float(g["length"]), G=int(g["G"]), dt=float(g["dt"]), total_steps=int(g["T"]), pbl=pbl, pbr=pbr, fbl=fbl, fbr=fbr, solver=g["solver"], box_yz=g["box_yz"])
    if g["cn"]:  # (one oracle only for the implicit stepper: this guards it against drift)
        out = L.run_CN(g["x0"], g["v0"], g["q"], g["m"], g["qm"], **kw)
        assert out["picard_iterations"].tolist() == g["picard_iterations"].tolist()
    else:
        out = C.run(g["x0"], g["v0"], g["q"], g["m"], g["qm"], ext_E=g["ext_E"], ext_B=g["ext_B"], **kw)
    for k in FIELDS + ("positions", "velocities"):
        assert _relerr(out[k], g[k]) < 1e-10, k
    assert _relerr(out["fields"][0], g["E0"]) < (1e-10 if "refsrc_" in path else 1e-12)
    assert _relerr(out["initial_velocities"], g["initial_velocities"]) < 1e-15


# The GPU legs of the vectors added after the round's GPU budget was spent (every refsrc_* file and the five newest cases) live in
# tests/test_zz_refsrc_gpu.py, which sorts last: the driver runs `pytest -x`, and a surprise there must not hide the rest of the suite.
GPU_VERIFIED = ("two_stream_periodic", "large_cfl_jumps", "reflective_absorbing", "absorbing_reflective_nofilter", "weibel_external_B",
                "relativistic", "field_solver_gauss_fft", "field_solver_cartesian_reflective", "crank_nicolson_periodic", "crank_nicolson_absorbing")
GPU_FILES = [f for f in FILES if os.path.basename(f)[:-4] in GPU_VERIFIED]
GPU_FILES_LATE = [f for f in FILES if f not in GPU_FILES]


def _species(g):
    ne, ni = int(g["n_e"]), int(g["n_i"])
    return [dict(count=ne, q=float(g["q"][0]), m=float(g["m"][0]), qm=float(g["qm"][0])),
            dict(count=ni, q=float(g["q"][ne]), m=float(g["m"][ne]), qm=float(g["qm"][ne]))]


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["indexed", "binned"])
@pytest.mark.parametrize("path", GPU_FILES, ids=[os.path.basename(f)[:-4] for f in GPU_FILES])
def test_cuda_reproduces_golden(path, engine):
    cuda_reproduces_golden(path, engine)


def cuda_reproduces_golden(path, engine):
    import torch
    from jaxincell_b200 import HotPath
    g = _load(path)
    pbl, pbr, fbl, fbr = (int(b) for b in g["bcs"])
    s = g["solver"]
    cn = g["cn"]
    if cn and engine == "binned":
        pytest.skip("the Crank-Nicolson stepper has one particle store")
    ordered = engine == "indexed" or cn  # particle histories / initial velocities in input order
    extra = dict(time_evolution_algorithm=1, cn_substeps=s["number_of_particle_substeps_implicit_CN"],
                 cn_max_iterations=s["max_number_of_Picard_iterations_implicit_CN"], cn_tolerance=s["tolerance_Picard_iterations_implicit_CN"]) if cn else {}
    if g["box_yz"] is not None:
        extra.update(length_y=g["box_yz"][0], length_z=g["box_yz"][1])
    hp = HotPath(species=_species(g), length=float(g["length"]), G=int(g["G"]), dt=float(g["dt"]), pbl=pbl, pbr=pbr, fbl=fbl, fbr=fbr,
                 filter_passes=s["filter_passes"], filter_alpha=s["filter_alpha"], filter_strides=s["filter_strides"],
                 relativistic=s["relativistic"], engine=engine, track_yz=engine == "indexed", field_solver=s["field_solver"], **extra)
    hp.set_external_fields(g["ext_E"], g["ext_B"])
    hp.initialize(g["x0"], g["v0"])
    out = hp.run(int(g["T"]), particles=ordered)
    torch.cuda.synchronize()
    for k in FIELDS + (("positions", "velocities") if ordered else ()):
        assert _relerr(out[k].cpu().numpy(), g[k]) < RTOL_F64, (k, engine)
    E0, B0, vi = hp.initial(velocities=ordered)
    assert _relerr(E0.cpu().numpy(), g["E0"]) < RTOL_F64
    if ordered:
        assert _relerr(vi.cpu().numpy(), g["initial_velocities"]) < 1e-14
    if cn:
        assert hp.picard_iterations()[1] == int(g["picard_iterations"].sum())
    hp.close()
