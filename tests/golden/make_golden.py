#!/usr/bin/env python
"""Regenerate tests/golden/*.npz:  python tests/golden/make_golden.py

The vectors are produced by oracle/literal.py -- the expression-by-expression NumPy restatement of the reference.  Their twins
tests/golden/refsrc_<case>.npz (tests/golden/make_reference_golden.py) are produced from the same CASES table by the reference's own
source on the NumPy stand-in for jax (tests/refshim); the two families agree to round-off (tests/golden/REFERENCE_SOURCE_RUN.md).
Each file holds the inputs (x0, v0, per-particle q, m, q/m, geometry, BCs, solver switches, external fields) and the six
per-step histories of jaxincell/_algorithms.py:93 plus the initial fields / post-BC initial velocities."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

from oracle import literal as L  # noqa: E402
from plasma import cfl_dt, two_species  # noqa: E402

CASES = {
    # name: (two_species kwargs, G, length, cfl, T, bcs(pbl,pbr,fbl,fbr), solver overrides, external field amplitude)
    "two_stream_periodic": (dict(n_e=160, n_i=120, seed=5, vth_e=0.05, vth_yz=0.02, drift=6e7, plus_minus=True, gpdl=0.6), 16, 0.01, 0.9, 10, (0, 0, 0, 0), {}, 0.0),
    "large_cfl_jumps": (dict(n_e=120, n_i=80, seed=6, vth_e=0.05, drift=6e7, plus_minus=True, gpdl=0.6), 20, 0.01, 4.5, 8, (0, 0, 0, 0), {}, 0.0),
    "reflective_absorbing": (dict(n_e=140, n_i=100, seed=7, vth_e=0.08, vth_yz=0.03, gpdl=0.6), 12, 0.01, 1.3, 10, (1, 2, 1, 2), {}, 0.0),
    "absorbing_reflective_nofilter": (dict(n_e=100, n_i=100, seed=8, vth_e=0.08, vth_yz=0.03, gpdl=0.6), 9, 0.01, 1.1, 8, (2, 1, 2, 1), {"filter_passes": 0}, 0.0),
    "weibel_external_B": (dict(n_e=150, n_i=150, seed=9, vth_e=0.01, vth_yz=0.1, gpdl=0.6, ion_vth_scale=1.0), 14, 0.02, 1.0, 8, (0, 0, 0, 0), {"filter_passes": 2, "filter_strides": (1, 3)}, 1.0),
    "relativistic": (dict(n_e=100, n_i=60, seed=10, vth_e=0.3, vth_yz=0.2, gpdl=0.6), 10, 0.01, 0.7, 8, (0, 0, 0, 0), {"relativistic": True}, 0.3),
    # per-step electrostatic correction (_algorithms.py:69-78) and the implicit stepper (_algorithms.py:100-241)
    "field_solver_gauss_fft": (dict(n_e=160, n_i=120, seed=11, vth_e=0.05, vth_yz=0.02, drift=6e7, plus_minus=True, gpdl=0.6), 16, 0.01, 0.9, 10, (0, 0, 0, 0), {"field_solver": 1}, 0.0),
    "field_solver_cartesian_reflective": (dict(n_e=140, n_i=100, seed=12, vth_e=0.08, vth_yz=0.03, gpdl=0.6), 12, 0.01, 0.9, 10, (1, 1, 1, 1), {"field_solver": 2, "filter_passes": 2}, 0.0),
    "crank_nicolson_periodic": (dict(n_e=160, n_i=120, seed=13, vth_e=0.05, vth_yz=0.02, drift=4e7, plus_minus=True, gpdl=0.03), 16, 0.01, 0.3, 8, (0, 0, 0, 0),
                                {"time_evolution_algorithm": 1, "max_number_of_Picard_iterations_implicit_CN": 12, "number_of_particle_substeps_implicit_CN": 2,
                                 "tolerance_Picard_iterations_implicit_CN": 1e-9}, 0.0),
    # the remaining solver, a solver with walls, grids smaller than the 6-node J_x window (_sources.py:199-204), a box with
    # length_y, length_z of their own (transverse wrap of _boundary_conditions.py:7-145)
    "field_solver_poisson_fft": (dict(n_e=140, n_i=100, seed=15, vth_e=0.05, vth_yz=0.02, drift=5e7, plus_minus=True, gpdl=0.6), 14, 0.01, 0.8, 8, (0, 0, 0, 0), {"field_solver": 3}, 0.0),
    "field_solver_gauss_fft_absorbing": (dict(n_e=120, n_i=100, seed=16, vth_e=0.08, vth_yz=0.03, gpdl=0.6), 12, 0.01, 0.9, 8, (2, 2, 2, 2), {"field_solver": 1, "filter_passes": 1}, 0.0),
    "tiny_grid_5_cells": (dict(n_e=60, n_i=50, seed=17, vth_e=0.05, vth_yz=0.02, drift=4e7, plus_minus=True, gpdl=0.6), 5, 0.01, 1.8, 8, (0, 0, 0, 0), {"filter_strides": (1, 2)}, 0.0),
    "tiny_grid_3_cells_reflective": (dict(n_e=50, n_i=40, seed=18, vth_e=0.08, vth_yz=0.03, gpdl=0.6), 3, 0.01, 0.9, 8, (1, 1, 1, 1), {"filter_passes": 2, "filter_strides": (1,)}, 0.0),
    "transverse_box": (dict(n_e=100, n_i=80, seed=19, vth_e=0.05, vth_yz=0.2, gpdl=0.6, box_yz=(0.004, 0.02)), 12, 0.01, 1.0, 10, (0, 0, 0, 0), {"relativistic": True}, 0.2),
    "crank_nicolson_absorbing": (dict(n_e=140, n_i=100, seed=14, vth_e=0.2, vth_yz=0.1, gpdl=0.01), 12, 0.01, 0.3, 8, (2, 2, 2, 2),
                                 {"time_evolution_algorithm": 1, "max_number_of_Picard_iterations_implicit_CN": 5, "number_of_particle_substeps_implicit_CN": 3,
                                  "tolerance_Picard_iterations_implicit_CN": 1e-30}, 0.0),
}


def build(name):
    kw, G, length, cfl, T, bcs, solver, ext = CASES[name]
    kw = dict(kw)
    n_e, n_i = kw.pop("n_e"), kw.pop("n_i")
    box_yz = kw.pop("box_yz", None)
    p = two_species(n_e, n_i, length=length, G=G, **kw)
    if box_yz is not None:  # two_species draws y, z in the x extent: rescale into the transverse box
        p["x0"][:, 1] *= box_yz[0] / length
        p["x0"][:, 2] *= box_yz[1] / length
    dt = cfl_dt(length, G, cfl)
    rng = np.random.default_rng(99)
    ext_E = (ext * 1e3 * rng.standard_normal((G, 3))).astype(np.float32) if ext else None
    ext_B = (ext * 1e-3 * rng.standard_normal((G, 3))).astype(np.float32) if ext else None
    pbl, pbr, fbl, fbr = bcs
    cn = solver.get("time_evolution_algorithm", 0) == 1
    if cn:
        out = L.run_CN(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T, pbl=pbl, pbr=pbr, fbl=fbl,
                       fbr=fbr, solver=solver)
    else:
        out = L.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T, pbl=pbl, pbr=pbr, fbl=fbl,
                    fbr=fbr, solver=solver, ext_E=ext_E, ext_B=ext_B, box_yz=box_yz)
    sol = {"filter_passes": 5, "filter_alpha": 0.5, "filter_strides": (1, 2, 4), "relativistic": False, "field_solver": 0,
           "time_evolution_algorithm": 0, "max_number_of_Picard_iterations_implicit_CN": 20, "number_of_particle_substeps_implicit_CN": 2,
           "tolerance_Picard_iterations_implicit_CN": 1e-6, **solver}
    extra = dict(field_solver=sol["field_solver"], time_evolution_algorithm=sol["time_evolution_algorithm"],
                 cn_max_iterations=sol["max_number_of_Picard_iterations_implicit_CN"], cn_substeps=sol["number_of_particle_substeps_implicit_CN"],
                 cn_tolerance=sol["tolerance_Picard_iterations_implicit_CN"])
    if cn:
        extra["picard_iterations"] = out["picard_iterations"]
    if box_yz is not None:
        extra["box_yz"] = np.array(box_yz)
    return dict(x0=p["x0"], v0=p["v0"], q=p["q"], m=p["m"], qm=p["qm"], n_e=n_e, n_i=n_i, length=length, G=G, dt=dt, T=T,
                bcs=np.array(bcs), filter_passes=sol["filter_passes"], filter_alpha=sol["filter_alpha"],
                filter_strides=np.array(sol["filter_strides"]), relativistic=int(sol["relativistic"]),
                ext_E=np.zeros((G, 3), np.float32) if ext_E is None else ext_E, ext_B=np.zeros((G, 3), np.float32) if ext_B is None else ext_B,
                positions=out["positions"], velocities=out["velocities"], electric_field=out["electric_field"],
                magnetic_field=out["magnetic_field"], current_density=out["current_density"], charge_density=out["charge_density"],
                E0=out["fields"][0], B0=out["fields"][1], initial_velocities=out["initial_velocities"], **extra)


if __name__ == "__main__":
    for name in CASES:  # existing files are kept (pass --all to regenerate everything)
        path = os.path.join(HERE, name + ".npz")
        if os.path.exists(path) and "--all" not in sys.argv:
            continue
        np.savez_compressed(path, **build(name))
        print("wrote", name)
