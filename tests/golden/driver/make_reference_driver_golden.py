#!/usr/bin/env python
"""Host-driver golden values from the REFERENCE'S OWN SOURCE on the NumPy stand-in (tests/refshim):

    python tests/golden/driver/make_reference_driver_golden.py

Runs only where /root/reference exists (the build container).  TEST INFRASTRUCTURE.  For every parameter dict of
tests/golden/driver/driver_cases.py it builds the reference's `Simulation(parameters)` (jaxincell/_simulation.py:85-92: cleaners, species
cross references, domain / particle / field state) and records what the host side of the hot path must reproduce:

  * domain state: dx, dt, grid end points, box size (_state_initialization.py:27-49);
  * per species, in the reference's order: count, resolved vth / drift / amplitudes / flags, the (seed_position, seed_velocity) pair
    that actually reached `initialize_species_phase_space` (_state_initialization.py:87-96,126-136), weight, charge, mass, q/m
    (:172-185,259-261);
  * the assembled particle state: positions / velocities drawn by the reference's formulas (:51-85) on the Threefry restatement
    (the stand-in's jax.random = oracle/sampling.py), after the 0.99c clip (:263-264);
  * the same host-state record for the reference's own example inputs `examples/input.toml` and `examples/bump-on-tail.toml` (read from
    the reference checkout, not copied), checked in this container only, where the checkout exists;
  * for RUN_CASE: `Simulation.run()` -> key set of the output dictionary, plasma_frequency, time_array end points (_simulation.py:263-312),
    and `diagnostics(output)` (_diagnostics.py:8-147) -> energies, dominant frequency, species names; the run's inputs/outputs that
    `diagnostics` consumes are stored too so that jaxincell_b200.diagnostics can be fed the same arrays.

Output: tests/golden/driver/refsrc_driver.json (scalars, tables) and tests/golden/driver/refsrc_driver_arrays.npz."""
import copy
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
REF = os.environ.get("JIC_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(ROOT, "tests", "refshim"), REF, ROOT, os.path.join(ROOT, "tests"), HERE]

import jax  # noqa: E402  (the stand-in)
assert "standin" in jax.__version__
import jaxincell  # noqa: E402
import jaxincell._state_initialization as SI  # noqa: E402
from driver_cases import CASES, RUN_CASE, RUN_CASES, RUNTIME_INVALID, RUNTIME_OVERRIDES  # noqa: E402

AXES = ("x", "y", "z")


def f(v):
    return float(np.asarray(v))


def record(name, params):
    seeds = []
    inner = SI.initialize_species_phase_space

    def spy(species, seed_position, seed_velocity, number_particles, box_size):
        seeds.append((int(seed_position), int(seed_velocity), int(number_particles)))
        return inner(species, seed_position, seed_velocity, number_particles, box_size)
    SI.initialize_species_phase_space = spy
    try:
        sim = jaxincell.Simulation(copy.deepcopy(params))  # a dict, or the path of a TOML file (_simulation.py:85-90)
    finally:
        SI.initialize_species_phase_space = inner
    # the sixteen numbers of the reference's start-up summary (print_simulation_information -> jax.debug.print), captured at the call
    info = []
    inner_print = SI.jprint
    SI.jprint = lambda fmt, *a, **k: info.append([f(v) for v in a])
    try:
        solver_on = dict(sim.solver_parameters, print_info=True)
        particle_state = {"weights": sim.weights, "charge_electrons": sim.charge_electrons, "vth_electrons": sim.vth_electrons, "velocities": sim.velocities}
        SI.print_simulation_information(sim.domain_parameters, sim.species_parameters, sim.external_field_parameters, solver_on,
                                        sim.current_domain_state(), particle_state)
    finally:
        SI.jprint = inner_print
    sp_all = sim.species_parameters
    species, o = [], 0
    k = 0
    for kind in ("electrons", "ions"):
        for canon, sp in sp_all[kind].items():
            n = sp["number_pseudoparticles"]
            species.append(dict(kind=kind, canonical=canon, user_label=sp["user_label"], count=n,
                                seed_position=seeds[k][0], seed_velocity=seeds[k][1],
                                weight=f(sim.weights[o, 0]), charge=f(sim.charges[o, 0]), mass=f(sim.masses[o, 0]),
                                charge_to_mass=f(sim.charge_to_mass_ratios[o, 0]),
                                grid_points_per_Debye_length=f(sp["grid_points_per_Debye_length"]),
                                **{f"{key}_{a}": (bool(sp[f"{key}_{a}"]) if isinstance(sp[f"{key}_{a}"], bool) else f(sp[f"{key}_{a}"]))
                                   for key in ("vth_over_c", "drift_speed", "perturbation_amplitude", "perturbation_wavenumber",
                                               "random_positions", "velocity_plus_minus") for a in AXES}))
            assert seeds[k][2] == n
            o += n
            k += 1
    rec = dict(information=info[0], dx=f(sim.dx), dt=f(sim.dt), grid_first=f(sim.grid[0]), grid_last=f(sim.grid[-1]), grid_size=int(len(sim.grid)),
               box_size=[f(b) for b in sim.box_size], species=species, n_particles=int(o),
               solver={k: (list(v) if isinstance(v, tuple) else v) for k, v in sim.solver_parameters.items()},
               domain={k: (f(v) if not isinstance(v, (int, bool)) else v) for k, v in sim.domain_parameters.items()})
    arrays = {f"{name}__positions": np.asarray(sim.positions, np.float64), f"{name}__velocities": np.asarray(sim.velocities, np.float64),
              f"{name}__E0": np.asarray(sim.fields[0], np.float64)}
    return sim, rec, arrays


def reference_growth_rate(out):
    """`energy_gamma_from_output` of the reference's examples/inference_two_stream.py:108-203 (the growth-rate diagnostic BASELINE.md names),
    executed from the file where it lies: the function and its two window constants are cut out of the script by `ast` (the script
    itself runs simulations and plots at import) and run on the stand-in."""
    import ast
    import jax.numpy as jnp
    path = os.path.join(REF, "examples", "inference_two_stream.py")
    tree = ast.parse(open(path).read())
    keep = [n for n in tree.body if (isinstance(n, ast.FunctionDef) and n.name == "energy_gamma_from_output")
            or (isinstance(n, ast.Assign) and any(getattr(t, "id", "").startswith("FIT_FRAC") for t in n.targets))]
    assert len(keep) == 3, [type(n).__name__ for n in keep]
    ns = {"jnp": jnp}
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    res = ns["energy_gamma_from_output"](out, for_jit=True)
    return f(res["gamma_amp"]), int(res["i_start"]), int(res["i_end"])


def runtime_sections(sim, overrides):
    """What `_simulation` works with after `run(overrides)`: cleaned runtime input merged into the base sections, references re-resolved
    (jaxincell/_simulation.py:94-117,148-163)."""
    from jaxincell._parameters._sections import PARAMETER_SECTIONS
    from jaxincell._parameters._species_parameters import resolve_species_references
    from jaxincell._routing import build_runtime_parameter_sections
    cleaned = sim.clean_runtime_input_parameters(copy.deepcopy(overrides))
    base = {name: getattr(sim, meta["attribute"]) for name, meta in PARAMETER_SECTIONS.items()}
    sec = build_runtime_parameter_sections(base, cleaned)
    resolve_species_references(sec["species_parameters"])
    species = []
    for kind in ("electrons", "ions"):
        for canon, sp in sec["species_parameters"][kind].items():
            rec = dict(kind=kind, canonical=canon, weight=f(sp["weight"]), grid_points_per_Debye_length=f(sp["grid_points_per_Debye_length"]),
                       charge_over_elementary_charge=f(sp["charge_over_elementary_charge"]),
                       **{f"{key}_{a}": f(sp[f"{key}_{a}"]) for key in ("vth_over_c", "drift_speed") for a in AXES})
            if kind == "ions":
                rec["mass_over_proton_mass"] = f(sp["mass_over_proton_mass"])
            species.append(rec)
    return dict(length=f(sec["domain_parameters"]["length"]), cfl=f(sec["domain_parameters"]["timestep_over_spatialstep_times_c"]),
                filter_alpha=f(sec["solver_parameters"]["filter_alpha"]), species=species)


TOML_CASES = ("examples/input.toml", "examples/bump-on-tail.toml")  # the reference's own inputs, read where they lie (never copied)


def main():
    out, arrays = {}, {}
    for rel in TOML_CASES:
        _, rec, _ = record(rel, os.path.join(REF, rel))
        out["toml:" + rel] = rec
        print(rel, rec["n_particles"], "particles,", len(rec["species"]), "species")
    for name, params in CASES.items():
        sim, rec, arr = record(name, params)
        out[name] = rec
        arrays.update(arr)
        print(name, rec["n_particles"], "particles,", len(rec["species"]), "species")
        if name in RUNTIME_OVERRIDES:
            rec["runtime_overrides"] = runtime_sections(sim, RUNTIME_OVERRIDES[name])
        rec["runtime_invalid"] = []
        for case, bad, exc, text in RUNTIME_INVALID:
            if case == name:
                try:
                    sim.clean_runtime_input_parameters(copy.deepcopy(bad))
                    rec["runtime_invalid"].append(["no error", ""])
                except Exception as e:  # noqa: BLE001
                    assert type(e).__name__ == exc and text in str(e), (bad, type(e).__name__, str(e))
                    rec["runtime_invalid"].append([type(e).__name__, str(e)])
        if name in RUN_CASES and name != RUN_CASE:
            res = sim.run()
            rec["plasma_frequency"] = f(res["plasma_frequency"])
            for k in ("positions", "velocities", "electric_field", "magnetic_field", "current_density", "charge_density", "initial_velocities"):
                arrays[f"{name}__run__{k}"] = np.asarray(res[k])
        if name == RUN_CASE:
            res = sim.run()
            rec["output_keys"] = sorted(k for k in res.keys())
            rec["plasma_frequency"] = f(res["plasma_frequency"])
            rec["time_array"] = [f(res["time_array"][0]), f(res["time_array"][1]), f(res["time_array"][-1]), int(len(res["time_array"]))]
            rec["max_initial_vth_electrons"] = f(res["max_initial_vth_electrons"])
            rec["number_pseudoelectrons"] = int(res["number_pseudoelectrons"])
            for k in ("positions", "velocities", "electric_field", "magnetic_field", "current_density", "charge_density", "masses", "charges",
                      "initial_velocities", "external_electric_field", "external_magnetic_field", "grid"):
                arrays[f"run__{k}"] = np.asarray(res[k])
            rec["growth_rate"], rec["growth_fit_start"], rec["growth_fit_end"] = reference_growth_rate(res)
            jaxincell.diagnostics(res)  # mutates `res` and returns None (_diagnostics.py:8-147)
            d = res
            rec["diagnostics_keys"] = sorted(d.keys())
            rec["species_names"] = [s["name"] for s in d["species"]]
            rec["dominant_frequency"] = f(d["dominant_frequency"])
            for k in ("electric_field_energy", "magnetic_field_energy", "kinetic_energy", "kinetic_energy_electrons", "kinetic_energy_ions",
                      "total_energy", "electric_field_energy_density", "external_electric_field_energy", "external_magnetic_field_energy"):
                arrays[f"diag__{k}"] = np.asarray(d[k], np.float64)
    with open(os.path.join(HERE, "refsrc_driver.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True, default=lambda v: f(v) if np.ndim(v) == 0 else np.asarray(v).tolist())
    np.savez_compressed(os.path.join(HERE, "refsrc_driver_arrays.npz"), **arrays)


if __name__ == "__main__":
    main()
