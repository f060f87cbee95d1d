#!/usr/bin/env python
"""Golden vectors from the REFERENCE'S OWN SOURCE, executed on the NumPy stand-in for jax (tests/refshim):

    python tests/golden/make_reference_golden.py            # unit tests of the reference + refsrc_*.npz
    python tests/golden/make_reference_golden.py --no-unit  # vectors only

Runs only where /root/reference exists (the build container); the .npz files it writes are what travels.  TEST INFRASTRUCTURE.

What it does
  1. runs the reference's own pytest files on the stand-in (everything except plotting and autodiff), which pins the stand-in
     to the reference's known-answer vectors, and records the outcome in tests/golden/REFERENCE_SOURCE_RUN.md;
  2. for every case of tests/golden/make_golden.py (same seeded particles, geometry, BCs, solver switches, external fields) builds
     the reference's parameter dict -- explicit per-species `initial_positions` / `initial_velocities` / `weight`
     (jaxincell/_parameters/_species_definitions.py:57-58, _state_initialization.py:147-150,185) -- and calls the reference's
     `Simulation(parameters).run()` (jaxincell/_simulation.py:94-121): start-up, initial field solve, `Boris_step` / `CN_step`
     under `lax.scan`, output dict -- all of it the reference's code, none of it ours;
  3. `field_solver` 2 and 3 are rejected by the reference's parameter cleaner (_parameters/_solver_parameters.py:45) although
     `Boris_step` implements them (_algorithms.py:69-78): those cases scan the reference's `Boris_step` directly from the
     reference's own start-up state, wired as in _simulation.py:216-257;
  4. writes tests/golden/refsrc_<case>.npz with the same keys as the oracle-made files, plus the Picard iteration counts of the CN
     cases (obtained by counting calls of the reference's `current_density_periodic_CN`, one per sub-step).

The stand-in is not JAX (see tests/refshim/README.md): what these vectors pin is the reference's *composition* -- call order,
argument wiring, which positions/velocities/charges enter which deposit, start-up, scan outputs -- on top of callee arithmetic that
the reference's own unit tests pin.  Floating-point differences to real XLA (FMA contraction, reduction order) are at round-off.
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("JIC_REFERENCE", "/root/reference")
SHIM = os.path.join(ROOT, "tests", "refshim")

UNIT_FILES = ["test_particles.py", "test_sources.py", "test_boundary_conditions.py", "test_fields.py", "test_filters.py", "test_constants.py",
              "test_algorithms.py", "test_simulation.py", "test_state_initialization.py", "test_diagnostics.py", "test_routing.py",
              "test_domain_parameters.py", "test_external_field_parameters.py", "test_solver_parameters.py", "test_source_parameters.py",
              "test_species_definitions.py", "test_species_parameters.py", "test_parameter_sections.py", "test_parameter_utils.py",
              "test_runtime_input_parameters.py", "test_main.py", "test_package_exports.py"]
EXPECTED_UNIT_FAILURES = {  # jax.grad is outside the stand-in (autodiff is out of scope); version.py is written by the build backend
    "test_fields.py::test_field_functions_are_differentiable_for_small_inputs",
    "test_filters.py::test_filter_scalar_field_jit_and_grad_compatible",
    "test_filters.py::test_filter_vector_field_jit_and_grad_compatible",
    "test_particles.py::test_particle_helpers_are_differentiable_for_small_inputs",
    "test_package_exports.py::test_version_metadata_is_importable",
}


def run_reference_unit_tests():
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([SHIM, REF]))
    cmd = [sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "-rf"] + [os.path.join(REF, "tests", f) for f in UNIT_FILES]
    res = subprocess.run(cmd, cwd="/tmp", env=env, capture_output=True, text=True)
    lines = res.stdout.strip().splitlines()
    failed = sorted(ln.split("tests/")[-1].split(" - ")[0] for ln in lines if ln.startswith("FAILED"))
    return lines[-1], failed


def import_reference():
    sys.path[:0] = [SHIM, REF, ROOT, os.path.join(ROOT, "tests"), HERE]
    import jax  # noqa: F401  (the stand-in)
    assert "standin" in jax.__version__
    import jaxincell
    return jaxincell


def reference_parameters(case):
    """make_golden.CASES entry -> the reference's nested parameter dict."""
    import make_golden as MG
    from oracle import literal as L
    from plasma import two_species
    kw, G, length, cfl, T, bcs, solver, ext = MG.CASES[case]
    kw = dict(kw)
    n_e, n_i = kw.pop("n_e"), kw.pop("n_i")
    box_yz = kw.pop("box_yz", None)
    p = two_species(n_e, n_i, length=length, G=G, **kw)
    if box_yz is not None:
        p["x0"][:, 1] *= box_yz[0] / length
        p["x0"][:, 2] *= box_yz[1] / length
    rng = np.random.default_rng(99)
    ext_E = (ext * 1e3 * rng.standard_normal((G, 3))).astype(np.float32) if ext else None
    ext_B = (ext * 1e-3 * rng.standard_normal((G, 3))).astype(np.float32) if ext else None
    we = p["q"][0] / -L.elementary_charge
    wi = p["q"][n_e] / L.elementary_charge
    ion_mass = kw.get("ion_mass", 1.0)
    pbl, pbr, fbl, fbr = bcs
    solver = dict(solver)
    field_solver = solver.pop("field_solver", 0)
    params = {
        "domain_parameters": dict(total_steps=T, timestep_over_spatialstep_times_c=cfl, number_grid_points=G, length=length,
                                  particle_BC_left=pbl, particle_BC_right=pbr, field_BC_left=fbl, field_BC_right=fbr),
        "solver_parameters": dict(print_info=False, field_solver=field_solver if field_solver in (0, 1) else 0, **solver),
        "species_parameters": {
            "electrons": {"electrons0": dict(number_pseudoparticles=n_e, weight=we, vth_over_c_x=kw.get("vth_e", 0.05),
                                             initial_positions=p["x0"][:n_e], initial_velocities=p["v0"][:n_e])},
            "ions": {"ions0": dict(number_pseudoparticles=n_i, weight=wi, mass_over_proton_mass=ion_mass,
                                   vth_over_c_x=0.0, vth_over_c_y=0.0, vth_over_c_z=0.0,
                                   initial_positions=p["x0"][n_e:], initial_velocities=p["v0"][n_e:])},
        },
        "external_field_parameters": {},
    }
    if box_yz is not None:
        params["domain_parameters"].update(length_y=box_yz[0], length_z=box_yz[1])
    if ext:
        params["external_field_parameters"] = {"external_electric_field": {"E": ext_E}, "external_magnetic_field": {"B": ext_B}}
    return params, p, field_solver, (ext_E, ext_B)


def scan_boris_directly(jaxincell, sim, field_solver):
    """_simulation.py:216-257 with the reference's functions, for the field_solver values its parameter cleaner refuses."""
    from jax import lax, numpy as jnp
    from jaxincell._algorithms import Boris_step
    from jaxincell._boundary_conditions import set_BC_particles, set_BC_positions
    dom, sol = sim.domain_parameters, dict(sim.solver_parameters)
    dx, dt, grid, box = sim.dx, sim.dt, sim.grid, sim.box_size
    pbl, pbr, fbl, fbr = (dom[k] for k in ("particle_BC_left", "particle_BC_right", "field_BC_left", "field_BC_right"))
    E, B = sim.fields
    x, v = sim.positions, sim.velocities
    x_plus, v, qs, ms, q_ms = set_BC_particles(x + (dt / 2) * v, v, sim.charges, sim.masses, sim.charge_to_mass_ratios, dx, grid, *box, pbl, pbr)
    x_minus = set_BC_positions(x - (dt / 2) * v, sim.charges, dx, grid, *box, pbl, pbr)
    ext = {"external_electric_field": sim.external_electric_field, "external_magnetic_field": sim.external_magnetic_field}
    carry = (E, B, x_minus, x, x_plus, v, qs, ms, q_ms)
    step = lambda c, i: Boris_step(c, i, sol, ext, dx, dt, grid, box, pbl, pbr, fbl, fbr, field_solver)  # noqa: E731
    _, res = lax.scan(step, carry, jnp.arange(dom["total_steps"]))
    names = ("positions", "velocities", "electric_field", "magnetic_field", "current_density", "charge_density")
    out = dict(zip(names, res))
    out.update(initial_velocities=v, fields=(E, B), dt=dt)
    return out


def run_case(jaxincell, case):
    params, p, field_solver, (ext_E, ext_B) = reference_parameters(case)
    sim = jaxincell.Simulation(params)
    cn = params["solver_parameters"].get("time_evolution_algorithm", 0) == 1
    picard = None
    if field_solver in (0, 1):
        if cn:  # one call of current_density_periodic_CN per sub-step of every Picard iteration (_algorithms.py:174-177)
            import jaxincell._algorithms as A
            calls, inner = [], A.current_density_periodic_CN
            A.current_density_periodic_CN = lambda *a, **k: (calls.append(1), inner(*a, **k))[1]
            out = sim.run()
            A.current_density_periodic_CN = inner
            sub = params["solver_parameters"]["number_of_particle_substeps_implicit_CN"]
            # the scan runs the steps in order, so the per-step counts are recovered from a second pass, step by step
            picard = count_picard_per_step(jaxincell, params, sub)
            assert sum(picard) * sub == len(calls), (sum(picard), sub, len(calls))
        else:
            out = sim.run()
    else:
        out = scan_boris_directly(jaxincell, sim, field_solver)
    G = params["domain_parameters"]["number_grid_points"]
    sol = {"filter_passes": 5, "filter_alpha": 0.5, "filter_strides": (1, 2, 4), "relativistic": False, "time_evolution_algorithm": 0,
           "max_number_of_Picard_iterations_implicit_CN": 20, "number_of_particle_substeps_implicit_CN": 2,
           "tolerance_Picard_iterations_implicit_CN": 1e-6, **params["solver_parameters"]}
    dom = params["domain_parameters"]
    a = lambda v: np.asarray(v, dtype=np.float64)  # noqa: E731
    rec = dict(x0=p["x0"], v0=p["v0"], q=p["q"], m=p["m"], qm=p["qm"], n_e=len(params["species_parameters"]["electrons"]["electrons0"]["initial_positions"]),
               n_i=len(params["species_parameters"]["ions"]["ions0"]["initial_positions"]), length=dom["length"], G=G, dt=float(out["dt"]),
               T=dom["total_steps"], bcs=np.array([dom["particle_BC_left"], dom["particle_BC_right"], dom["field_BC_left"], dom["field_BC_right"]]),
               filter_passes=sol["filter_passes"], filter_alpha=sol["filter_alpha"], filter_strides=np.array(sol["filter_strides"]),
               relativistic=int(sol["relativistic"]), ext_E=np.zeros((G, 3), np.float32) if ext_E is None else ext_E,
               ext_B=np.zeros((G, 3), np.float32) if ext_B is None else ext_B,
               positions=a(out["positions"]), velocities=a(out["velocities"]), electric_field=a(out["electric_field"]),
               magnetic_field=a(out["magnetic_field"]), current_density=a(out["current_density"]), charge_density=a(out["charge_density"]),
               E0=a(out["fields"][0]), B0=a(out["fields"][1]), initial_velocities=a(out["initial_velocities"]),
               field_solver=field_solver, time_evolution_algorithm=sol["time_evolution_algorithm"],
               cn_max_iterations=sol["max_number_of_Picard_iterations_implicit_CN"], cn_substeps=sol["number_of_particle_substeps_implicit_CN"],
               cn_tolerance=sol["tolerance_Picard_iterations_implicit_CN"], produced_by="reference source on tests/refshim")
    # the reference must have seen exactly the particles of the oracle-made case
    assert np.array_equal(a(sim.charges)[:, 0], p["q"]) or np.allclose(a(sim.charges)[:, 0], p["q"], rtol=1e-15, atol=0)
    assert np.allclose(a(sim.masses)[:, 0], p["m"], rtol=1e-15, atol=0) and np.allclose(a(sim.charge_to_mass_ratios)[:, 0], p["qm"], rtol=1e-15, atol=0)
    if picard is not None:
        rec["picard_iterations"] = np.array(picard)
    if float(dom.get("length_y", 0)):
        rec["box_yz"] = np.array([dom["length_y"], dom["length_z"]])
    return rec


def count_picard_per_step(jaxincell, params, sub):
    """Re-run with a counter that is read between the steps of the outer scan."""
    import jaxincell._algorithms as A
    import jaxincell._simulation as S
    calls, per_step, inner, inner_step = [0], [], A.current_density_periodic_CN, S.CN_step

    def counting(*a, **k):
        calls[0] += 1
        return inner(*a, **k)

    def step(*a, **k):
        before = calls[0]
        r = inner_step(*a, **k)
        per_step.append((calls[0] - before) // sub)
        return r
    A.current_density_periodic_CN, S.CN_step = counting, step
    try:
        jaxincell.Simulation(params).run()
    finally:
        A.current_density_periodic_CN, S.CN_step = inner, inner_step
    return per_step


def main():
    report = ["# Reference source executed on the NumPy stand-in (`tests/refshim`)", "",
              "Written by `tests/golden/make_reference_golden.py` in the build container (where `/root/reference` exists).", ""]
    if "--no-unit" not in sys.argv:
        summary, failed = run_reference_unit_tests()
        unexpected = [f for f in failed if f not in EXPECTED_UNIT_FAILURES]
        report += ["## The reference's own unit tests on the stand-in", "", f"`{summary.strip('= ')}`", "",
                   "Files: " + ", ".join(f"`tests/{f}`" for f in UNIT_FILES) + " (not run: `test_plots.py`, `test_autodifferentiability.py`).", "",
                   "Failures (all expected: `jax.grad` is outside the stand-in, `jaxincell/version.py` is written by the build backend):", ""]
        report += [f"* `{f}`" for f in failed] + [""]
        assert not unexpected, unexpected
        print(summary)
    jaxincell = import_reference()
    import make_golden as MG
    report += ["## Composed runs: reference `Simulation.run()` vs `oracle/literal.py` (max |difference| / max |reference|)", "",
               "| case | E | B | J | rho | x | v | Picard counts |", "|---|---|---|---|---|---|---|---|"]
    for case in MG.CASES:
        path = os.path.join(HERE, f"refsrc_{case}.npz")
        if os.path.exists(path) and "--all" not in sys.argv:  # existing files are kept (pass --all to regenerate everything)
            rec = dict(np.load(path))
        else:
            rec = run_case(jaxincell, case)
            np.savez_compressed(path, **rec)
        old = np.load(os.path.join(HERE, case + ".npz"))
        row = []
        for k in ("electric_field", "magnetic_field", "current_density", "charge_density", "positions", "velocities"):
            scale = max(np.abs(rec[k]).max(), 1e-300)
            row.append(f"{np.abs(rec[k] - old[k]).max() / scale:.1e}")
        pic = "—"
        if "picard_iterations" in rec:
            pic = "equal" if np.array_equal(rec["picard_iterations"], old["picard_iterations"]) else f"{rec['picard_iterations']} vs {old['picard_iterations']}"
        report.append(f"| `{case}` | " + " | ".join(row) + f" | {pic} |")
        print(report[-1])
    with open(os.path.join(HERE, "REFERENCE_SOURCE_RUN.md"), "w") as f:
        f.write("\n".join(report) + "\n")


if __name__ == "__main__":
    main()
