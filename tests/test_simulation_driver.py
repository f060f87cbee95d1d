"""The Simulation-shaped host driver (jaxincell_b200/_simulation.py) against the reference's public contract:
jaxincell/_simulation.py:43-344 and tests/test_simulation.py:79-144,246-262,668-721 of the reference, re-expressed without JAX."""
import numpy as np
import pytest

from jaxincell_b200 import JicError, Simulation, diagnostics, load_parameters
from jaxincell_b200 import _simulation as S
from oracle import closed_form as C

SMALL = {
    "domain_parameters": {"number_grid_points": 16, "total_steps": 12, "length": 0.01, "timestep_over_spatialstep_times_c": 0.9},
    "species_parameters": {
        "electrons": {"e": {"number_pseudoparticles": 300, "vth_over_c_x": 0.05, "vth_over_c_y": 0.02, "drift_speed_x": 5e7,
                            "velocity_plus_minus_x": True, "random_positions_x": True, "grid_points_per_Debye_length": 0.6}},
        "ions": {"i": {"number_pseudoparticles": 200, "vth_over_c_x": "e", "vth_over_c_y": "e", "vth_over_c_z": "e",
                       "random_positions_x": True, "grid_points_per_Debye_length": 0.6}},
    },
    "solver_parameters": {"print_info": False},
}


def test_defaults_match_the_reference_tables():
    sim = Simulation()
    assert sim.domain_parameters["number_grid_points"] == 50 and sim.domain_parameters["total_steps"] == 350  # _domain_parameters.py:12-25
    assert sim.solver_parameters["filter_strides"] == (1, 2, 4) and sim.solver_parameters["seed"] == 1701        # _solver_parameters.py:10-22
    e0 = sim.species_parameters["electrons"]["_electrons0"]
    assert e0["number_pseudoparticles"] == 500 and e0["drift_speed_x"] == 1e8 and e0["velocity_plus_minus_x"]  # _species_definitions.py:86-101
    i0 = sim.species_parameters["ions"]["_ions0"]
    # "_electrons0" reference: sqrt(T_i/T_e) * vth_e * sqrt(m_e / m_i)   (_species_parameters.py:112-121)
    np.testing.assert_allclose(i0["vth_over_c_x"], 0.05 * np.sqrt(S.mass_electron / S.mass_proton), rtol=1e-15)
    assert i0["vth_over_c_y"] == 0


def test_species_cross_references_and_validation():
    sim = Simulation(SMALL)
    i = sim.species_parameters["ions"]["_ions0"]
    np.testing.assert_allclose(i["vth_over_c_y"], 0.02 * np.sqrt(S.mass_electron / S.mass_proton), rtol=1e-15)
    assert i["user_label"] == "i"
    bad = {"species_parameters": {"ions": {"vth_over_c_x": "nobody"}}}
    with pytest.raises(ValueError):
        Simulation(bad)
    with pytest.raises(AssertionError):
        Simulation({"domain_parameters": {"total_steps": 0}})
    with pytest.raises(AssertionError):
        Simulation({"solver_parameters": {"filter_alpha": 1.5}})
    with pytest.raises(AssertionError):
        Simulation({"species_parameters": {"electrons": {"number_pseudoparticles": 0}}})


def test_particle_state_shapes_weights_and_seed_schedule():
    sim = Simulation({**SMALL, "solver_parameters": {"print_info": False, "rng": "numpy"}})  # host streams: runs without a GPU
    st = sim.build_domain_state(sim.domain_parameters)
    np.testing.assert_allclose(st["dt"], 0.9 * st["dx"] / S.speed_of_light)
    np.testing.assert_allclose(st["grid"][0], -0.005 + st["dx"] / 2)
    ps = sim.initialize_particle_state(sim.species_parameters, sim.domain_parameters, sim.solver_parameters, st)
    N = 500
    assert ps["positions"].shape == (N, 3) and ps["velocities"].shape == (N, 3) and ps["charges"].shape == (N, 1)
    assert (np.abs(ps["positions"]) <= 0.005 + 1e-12).all()
    # auto weight (_state_initialization.py:172-185)
    w = S.epsilon_0 * S.mass_electron * S.speed_of_light ** 2 / S.elementary_charge ** 2 * 16 ** 2 / 0.01 / (2 * 300) * 0.05 ** 2 * 0.6 ** 2
    np.testing.assert_allclose(ps["weights"][0, 0], w, rtol=1e-14)
    np.testing.assert_allclose(ps["charges"][0, 0], -S.elementary_charge * w, rtol=1e-14)
    np.testing.assert_allclose(ps["charge_to_mass_ratios"][-1, 0], S.elementary_charge / S.mass_proton, rtol=1e-14)
    # alternating drift sign, thermal spread of the right size
    vx = ps["velocities"][:300, 0]
    assert (np.sign(vx[::2]) > 0).mean() > 0.95 and (np.sign(vx[1::2]) < 0).mean() > 0.95
    assert abs(np.std(np.abs(vx)) / (0.05 * S.speed_of_light / np.sqrt(2)) - 1) < 0.2
    again = sim.initialize_particle_state(sim.species_parameters, sim.domain_parameters, sim.solver_parameters, st)
    np.testing.assert_array_equal(ps["positions"], again["positions"])
    assert S._seed_pair(1701, "electrons", 0, 0) == (1701, 1704) and S._seed_pair(1701, "ions", 0, 0) == (1701, 1707)
    assert S._seed_pair(1701, "electrons", 1, 0) == (1713, 1713) and S._seed_pair(1701, "ions", 1, 1) == (1719, 1719)


def test_load_parameters_toml(tmp_path):
    f = tmp_path / "input.toml"
    f.write_text('[domain_parameters]\nlength = 0.01\nnumber_grid_points = 70\ntotal_steps = 5\n'
                 '[solver_parameters]\nfilter_strides = [1, 2, 4]\nprint_info = false\n'
                 '[species_parameters.electrons.electrons0]\nnumber_pseudoparticles = 40\nvth_over_c_x = 0.05\n'
                 '[species_parameters.ions.ions0]\nnumber_pseudoparticles = 40\nvth_over_c_x = "_electrons0"\n')
    p = load_parameters(str(f))
    assert p["domain_parameters"]["number_grid_points"] == 70
    sim = Simulation(str(f))
    assert sim.solver_parameters["filter_strides"] == (1, 2, 4)
    assert sim.species_parameters["electrons"]["_electrons0"]["number_pseudoparticles"] == 40


def test_out_of_scope_algorithms_are_rejected_loudly():
    with pytest.raises(AssertionError):  # _solver_parameters.py:44
        Simulation({**SMALL, "solver_parameters": {"print_info": False, "time_evolution_algorithm": 2}})
    with pytest.raises(AssertionError):  # _solver_parameters.py:43 only admits 0 and 1
        Simulation({**SMALL, "solver_parameters": {"print_info": False, "field_solver": 2}})


def test_diagnostics_energies_on_a_fabricated_output():
    T, N, G = 6, 5, 8
    rng = np.random.default_rng(0)
    q = np.array([-1.0, -1.0, 2.0, 2.0, 2.0]).reshape(-1, 1); m = np.array([1.0, 1.0, 3.0, 3.0, 3.0]).reshape(-1, 1)
    out = dict(positions=rng.normal(size=(T, N, 3)), velocities=rng.normal(size=(T, N, 3)), charges=q, masses=m,
               electric_field=rng.normal(size=(T, G, 3)), magnetic_field=rng.normal(size=(T, G, 3)),
               external_electric_field=np.zeros((G, 3), np.float32), external_magnetic_field=np.ones((G, 3), np.float32),
               grid=np.arange(G) * 0.5, total_steps=T, dt=0.1, dx=0.5, plasma_frequency=1.0)
    v = out["velocities"].copy(); E = out["electric_field"].copy()
    d = diagnostics(out)
    assert "positions" not in d and d["position_electrons"].shape == (T, 2, 3) and d["velocity_ions"].shape == (T, 3, 3)
    np.testing.assert_allclose(d["kinetic_energy"], 0.5 * np.sum(m.reshape(-1) * np.sum(v ** 2, axis=-1), axis=-1))
    np.testing.assert_allclose(d["electric_field_energy"], S.epsilon_0 / 2 * np.sum(E ** 2, axis=(1, 2)) * 0.5)
    np.testing.assert_allclose(d["external_magnetic_field_energy"], 1 / (2 * S.mu_0) * 3 * G * 0.5)
    assert [s["name"] for s in d["species"]] == ["electrons", "ions"]
    assert d["total_energy"].shape == (T,)


# ---- against the reference's own source (tests/golden/driver: written by make_reference_driver_golden.py on tests/refshim) -----------
import copy  # noqa: E402
import json  # noqa: E402
import os  # noqa: E402
import sys  # noqa: E402

_DRV = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "driver")
sys.path.insert(0, _DRV)
from driver_cases import CASES as DRIVER_CASES, RUN_CASE, RUN_CASES, RUNTIME_INVALID, RUNTIME_OVERRIDES  # noqa: E402

with open(os.path.join(_DRV, "refsrc_driver.json")) as _f:
    REFSRC = json.load(_f)


REFERENCE_CHECKOUT = os.environ.get("JIC_REFERENCE", "/root/reference")
TOML_CASES = [k for k in REFSRC if k.startswith("toml:")]  # the reference's own example inputs: read from its checkout, never copied


def _host_state(name, rng="numpy"):
    if name.startswith("toml:"):
        path = os.path.join(REFERENCE_CHECKOUT, name[len("toml:"):])
        if not os.path.exists(path):
            pytest.skip("the reference checkout exists in the build container only")
        sim = Simulation(path)
        sim.solver_parameters["rng"] = rng  # host streams: no GPU needed for the tables compared here
    else:
        par = copy.deepcopy(DRIVER_CASES[name])
        par.setdefault("solver_parameters", {})["rng"] = rng
        sim = Simulation(par)
    st = sim.build_domain_state(sim.domain_parameters)
    return sim, st


@pytest.mark.parametrize("name", sorted(DRIVER_CASES) + TOML_CASES)
def test_host_state_matches_the_reference_source(name):
    """Cleaners, species cross references, domain state, seed schedule, weights / charges / masses: equal to what the reference's
    `Simulation(parameters)` derives from the same dictionary (_simulation.py:85-92, _state_initialization.py:27-49,87-185,259-261)."""
    ref = REFSRC[name]
    sim, st = _host_state(name)
    np.testing.assert_allclose([st["dx"], st["dt"], st["grid"][0], st["grid"][-1]], [ref["dx"], ref["dt"], ref["grid_first"], ref["grid_last"]], rtol=1e-15)
    assert len(st["grid"]) == ref["grid_size"]
    np.testing.assert_allclose(st["box_size"], ref["box_size"], rtol=0)
    for k, v in ref["solver"].items():
        if k == "print_info" and name.startswith("toml:"):
            continue
        got = sim.solver_parameters[k]
        assert (list(got) if isinstance(got, tuple) else got) == v, k
    for k, v in ref["domain"].items():
        assert sim.domain_parameters[k] == v, k
    mine = [(kind, canon, sp) for kind in ("electrons", "ions") for canon, sp in sim.species_parameters[kind].items()]
    assert [(k, c, sp["user_label"], sp["number_pseudoparticles"]) for k, c, sp in mine] == \
        [(r["kind"], r["canonical"], r["user_label"], r["count"]) for r in ref["species"]]
    for (kind, canon, sp), r in zip(mine, ref["species"]):
        for key in ("vth_over_c", "drift_speed", "perturbation_amplitude", "perturbation_wavenumber"):
            for a in "xyz":
                np.testing.assert_allclose(float(sp[f"{key}_{a}"]), r[f"{key}_{a}"], rtol=1e-15, err_msg=f"{canon} {key}_{a}")
        for key in ("random_positions", "velocity_plus_minus"):
            for a in "xyz":
                assert bool(sp[f"{key}_{a}"]) == r[f"{key}_{a}"], (canon, key, a)
        np.testing.assert_allclose(float(sp["grid_points_per_Debye_length"]), r["grid_points_per_Debye_length"], rtol=1e-15)
    ps = sim.initialize_particle_state(sim.species_parameters, sim.domain_parameters, sim.solver_parameters, st)
    assert len(ps["positions"]) == ref["n_particles"]
    o = 0
    for r, t, smp in zip(ref["species"], ps["species_table"], ps["sampling"]):
        np.testing.assert_allclose([ps["weights"][o, 0], t["q"], t["m"], t["qm"]], [r["weight"], r["charge"], r["mass"], r["charge_to_mass"]], rtol=0)
        assert (smp["seed_position"], smp["seed_velocity"]) == (r["seed_position"], r["seed_velocity"]), r["canonical"]
        o += r["count"]


@pytest.mark.parametrize("name", sorted(DRIVER_CASES))
def test_initial_particles_of_the_reference_source_equal_the_sampling_oracle(name):
    """The reference's `initialize_species_phase_space` (_state_initialization.py:51-85) ran on the Threefry restatement
    (oracle/sampling.py as the stand-in's jax.random); oracle.sampling.sample composes the same draws itself (linspace form,
    perturbation, drift, (-1)^i, seed offsets +1..3 / +4..6) and the device kernel is tested against it (tests/test_sampling.py)."""
    from oracle import sampling as OS
    arrays = np.load(os.path.join(_DRV, "refsrc_driver_arrays.npz"))
    sim, st = _host_state(name)
    ps = sim.initialize_particle_state(sim.species_parameters, sim.domain_parameters, sim.solver_parameters, st)
    x, v = OS.sample(ps["sampling"], st["box_size"], partitionable=True)
    lim = 0.99 * S.speed_of_light
    v = np.where(np.abs(v) >= lim, np.sign(v) * lim, v)
    np.testing.assert_allclose(x, arrays[f"{name}__positions"], rtol=1e-15, atol=1e-18)
    np.testing.assert_allclose(v, arrays[f"{name}__velocities"], rtol=1e-15, atol=1e-9)


def _run_arrays(name):
    a = np.load(os.path.join(_DRV, "refsrc_driver_arrays.npz"))
    prefix = "run__" if name == RUN_CASE else f"{name}__run__"
    return a, {k[len(prefix):]: a[k] for k in a.files if k.startswith(prefix)}


@pytest.mark.parametrize("name", sorted(RUNTIME_OVERRIDES))
def test_runtime_overrides_resolve_like_the_reference_source(name):
    """`run(input_parameters)`: the sections the step works with after the overrides, against the reference's
    clean_runtime_input_parameters + build_runtime_parameter_sections + resolve_species_references (recorded by the generator).
    A changed electron thermal speed must reach the ions that reference "_electrons0"; a flat species value reaches every species of
    the type; canonical and user labels both address a species."""
    ref = REFSRC[name]["runtime_overrides"]
    sim, _ = _host_state(name)
    sec = sim._sections(sim.clean_runtime_input_parameters(copy.deepcopy(RUNTIME_OVERRIDES[name])))
    np.testing.assert_allclose([sec["domain_parameters"]["length"], sec["domain_parameters"]["timestep_over_spatialstep_times_c"],
                                sec["solver_parameters"]["filter_alpha"]], [ref["length"], ref["cfl"], ref["filter_alpha"]], rtol=0)
    mine = [(kind, canon, sp) for kind in ("electrons", "ions") for canon, sp in sec["species_parameters"][kind].items()]
    assert [(k, c) for k, c, _ in mine] == [(r["kind"], r["canonical"]) for r in ref["species"]]
    for (kind, canon, sp), r in zip(mine, ref["species"]):
        for key in r:
            if key in ("kind", "canonical"):
                continue
            np.testing.assert_allclose(float(sp[key]), r[key], rtol=1e-15, err_msg=f"{canon}.{key}")


def test_runtime_input_errors_match_the_reference_source():
    """Same exception types, and the offending path named, as the reference's clean_runtime_input_parameters (_routing.py:160-226)."""
    recorded = {name: list(REFSRC[name]["runtime_invalid"]) for name in REFSRC if REFSRC[name].get("runtime_invalid")}
    for case, bad, exc, text in RUNTIME_INVALID:
        ref_exc, ref_msg = recorded[case].pop(0)
        assert ref_exc == exc and text in ref_msg
        sim, _ = _host_state(case)
        with pytest.raises({"ValueError": ValueError, "TypeError": TypeError}[exc]) as err:
            sim.clean_runtime_input_parameters(copy.deepcopy(bad))
        assert text in str(err.value)


@pytest.mark.parametrize("name", sorted(DRIVER_CASES))
def test_state_attributes_of_the_simulation_object(name):
    """The reference's Simulation exposes its initial state as attributes (jaxincell/_simulation.py:438-492; used e.g. by its
    tests/test_simulation.py:640-700): dx, dt, grid, box_size, positions, velocities, weights, charges, masses, q/m."""
    ref = REFSRC[name]
    sim, st = _host_state(name)
    assert sim.dx == ref["dx"] and sim.dt == ref["dt"] and list(sim.box_size) == ref["box_size"]
    assert sim.grid.shape == (ref["grid_size"],) and sim.grid[0] == ref["grid_first"]
    N = ref["n_particles"]
    assert sim.positions.shape == (N, 3) and sim.velocities.shape == (N, 3)
    for attr in ("weights", "charges", "masses", "charge_to_mass_ratios"):
        assert getattr(sim, attr).shape == (N, 1), attr
    o = 0
    for r in ref["species"]:
        assert (sim.weights[o, 0], sim.charges[o, 0], sim.masses[o, 0], sim.charge_to_mass_ratios[o, 0]) == \
            (r["weight"], r["charge"], r["mass"], r["charge_to_mass"])
        o += r["count"]
    assert sim.positions is sim.positions  # cached
    with pytest.raises(AttributeError):
        sim.no_such_attribute


@pytest.mark.parametrize("name", sorted(DRIVER_CASES))
def test_start_up_summary_matches_the_reference_source(name, capsys):
    """The sixteen numbers `print_info` reports (print_simulation_information, _state_initialization.py:288-357), computed from the
    reference's own initial velocities (the gamma factors depend on the draws)."""
    ref = REFSRC[name]["information"]
    arrays = np.load(os.path.join(_DRV, "refsrc_driver_arrays.npz"))
    sim, st = _host_state(name)
    ps = sim.initialize_particle_state(sim.species_parameters, sim.domain_parameters, sim.solver_parameters, st)
    ps["velocities"] = arrays[f"{name}__velocities"]
    got = S.simulation_information(sim.domain_parameters, sim.species_parameters, sim.external_field_parameters, st, ps)
    np.testing.assert_allclose(np.array(got, dtype=float), np.array(ref, dtype=float), rtol=1e-13, atol=0)
    text = S.INFORMATION_TEXT.format(*got)
    assert text.count("\n") == 14 and "Skin Depths" in text and text.startswith("Length of the simulation box: ")


def test_growth_rate_diagnostic_matches_the_reference_source():
    """oracle.closed_form.growth_rate == `energy_gamma_from_output` of the reference's examples/inference_two_stream.py:108-203, which the
    generator cut out of that script and ran on the E_x history of the reference's own two-stream run (same window, same fit)."""
    ref = REFSRC[RUN_CASE]
    _, run = _run_arrays(RUN_CASE)
    T = ref["domain"]["total_steps"]
    assert (ref["growth_fit_start"], ref["growth_fit_end"]) == (int(0.30 * T), int(0.50 * T))
    got = C.growth_rate(run["electric_field"][:, :, 0], ref["dx"], ref["dt"], T)
    np.testing.assert_allclose(got, ref["growth_rate"], rtol=1e-9)


@pytest.mark.parametrize("name", RUN_CASES)
def test_oracle_reproduces_the_runs_of_the_reference_source(name):
    """The reference's `Simulation(parameters).run()` (on the stand-in, random particles from its own initialisation) against the
    closed-form oracle started from the same initial particles with the species table THIS driver derives from the same dictionary:
    five species with different weights, CFL 3 multi-cell jumps without filter / walls + relativistic / CFL 4.5 two-stream."""
    a, ref = _run_arrays(name)
    sim, st = _host_state(name)
    dom, sol = sim.domain_parameters, sim.solver_parameters
    ps = sim.initialize_particle_state(sim.species_parameters, dom, sol, st)
    out = C.run(a[f"{name}__positions"], a[f"{name}__velocities"], ps["charges"][:, 0], ps["masses"][:, 0], ps["charge_to_mass_ratios"][:, 0],
                length=st["box_size"][0], box_yz=st["box_size"][1:], G=int(dom["number_grid_points"]), dt=st["dt"], total_steps=int(dom["total_steps"]),
                pbl=dom["particle_BC_left"], pbr=dom["particle_BC_right"], fbl=dom["field_BC_left"], fbr=dom["field_BC_right"],
                solver=dict(filter_passes=sol["filter_passes"], filter_alpha=sol["filter_alpha"], filter_strides=sol["filter_strides"],
                            relativistic=sol["relativistic"]))
    for k in ("electric_field", "magnetic_field", "current_density", "charge_density", "positions", "velocities"):
        err = np.abs(out[k] - ref[k]).max() / max(np.abs(ref[k]).max(), 1e-300)
        assert err < 1e-9, (k, err)
    np.testing.assert_allclose(out["initial_velocities"], ref["initial_velocities"], rtol=1e-15)


def test_output_dictionary_keys_of_the_reference_source():
    """Every key of the reference's output dictionary (_simulation.py:269-344) is produced by the driver (checked on the key list the
    driver assembles; the values need a GPU run)."""
    ref = REFSRC[RUN_CASE]
    sim, st = _host_state(RUN_CASE)
    assert set(ref["output_keys"]) <= set(sim.output_keys()), sorted(set(ref["output_keys"]) - set(sim.output_keys()))


def test_diagnostics_match_the_reference_source():
    """jaxincell_b200.diagnostics on the arrays of the reference's run == the reference's own `diagnostics` (_diagnostics.py:8-147)."""
    ref = REFSRC[RUN_CASE]
    a = np.load(os.path.join(_DRV, "refsrc_driver_arrays.npz"))
    out = {k[len("run__"):]: a[k] for k in a.files if k.startswith("run__")}
    out.update(total_steps=ref["domain"]["total_steps"], dt=ref["dt"], dx=ref["dx"], plasma_frequency=ref["plasma_frequency"])
    d = diagnostics(out)
    assert [s["name"] for s in d["species"]] == ref["species_names"]
    np.testing.assert_allclose(d["dominant_frequency"], ref["dominant_frequency"], rtol=1e-12)
    for k in a.files:
        if k.startswith("diag__"):
            np.testing.assert_allclose(d[k[len("diag__"):]], a[k], rtol=1e-12, err_msg=k)
    missing = set(ref["diagnostics_keys"]) - set(d) - set(ref["output_keys"])
    assert not missing, sorted(missing)


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_output_contract_and_determinism():
    out = Simulation(SMALL).run()
    T, G, N = 12, 16, 500
    for k, shp in (("positions", (T, N, 3)), ("velocities", (T, N, 3)), ("masses", (N, 1)), ("charges", (N, 1)), ("weights", (N, 1)),
                   ("initial_positions", (N, 3)), ("initial_velocities", (N, 3)), ("species_integer_index", (N,)),
                   ("electric_field", (T, G, 3)), ("magnetic_field", (T, G, 3)), ("current_density", (T, G, 3)), ("charge_density", (T, G)),
                   ("grid", (G,)), ("time_array", (T,)), ("external_electric_field", (G, 3)), ("external_magnetic_field", (G, 3))):
        assert np.asarray(out[k]).shape == shp, k
    assert out["fields"][0].shape == (G, 3) and out["fields"][1].shape == (G, 3)
    assert out["number_grid_points"] == G and out["total_steps"] == T and out["plasma_frequency"] > 0
    assert all(np.isfinite(out[k]).all() for k in ("positions", "velocities", "electric_field", "charge_density"))
    assert (np.abs(out["positions"][..., 0]) <= 0.005).all()
    for k in ("domain_parameters", "species_parameters", "solver_parameters", "parameter_sections", "length", "filter_passes"):
        assert k in out
    again = Simulation(SMALL).run()
    np.testing.assert_allclose(again["electric_field"], out["electric_field"], rtol=1e-9, atol=1e-9 * np.abs(out["electric_field"]).max())
    d = diagnostics(out)
    drift = np.abs(d["total_energy"] - d["total_energy"][0]).max() / d["total_energy"][0]
    assert drift < 0.05


@pytest.mark.gpu
def test_explicit_initial_conditions_match_the_oracle():
    """The parity route that needs no RNG compatibility: initial_positions / initial_velocities per species
    (reference tests/test_simulation.py:668-721)."""
    rng = np.random.default_rng(3)
    ne, ni, G, L = 200, 150, 12, 0.01
    xe = rng.uniform(-L / 2, L / 2, (ne, 3)); xi = rng.uniform(-L / 2, L / 2, (ni, 3))
    ve = 1e7 * rng.standard_normal((ne, 3)); vi = 1e4 * rng.standard_normal((ni, 3))
    par = {"domain_parameters": {"number_grid_points": G, "total_steps": 15, "length": L, "timestep_over_spatialstep_times_c": 0.8,
                                 "particle_BC_left": 1, "particle_BC_right": 2, "field_BC_left": 1, "field_BC_right": 2},
           "species_parameters": {"electrons": {"number_pseudoparticles": ne, "vth_over_c_x": 0.05, "initial_positions": xe, "initial_velocities": ve,
                                                "grid_points_per_Debye_length": 0.6},
                                  "ions": {"number_pseudoparticles": ni, "initial_positions": xi, "initial_velocities": vi,
                                           "grid_points_per_Debye_length": 0.6}},
           "solver_parameters": {"print_info": False, "filter_passes": 3, "filter_strides": [1, 2]}}
    out = Simulation(par).run()
    dt = 0.8 * (L / G) / S.speed_of_light
    ref = C.run(np.concatenate([xe, xi]), np.concatenate([ve, vi]), out["charges"][:, 0], out["masses"][:, 0], out["charge_to_mass_ratios"][:, 0],
                length=L, G=G, dt=dt, total_steps=15, pbl=1, pbr=2, fbl=1, fbr=2,
                solver=dict(filter_passes=3, filter_alpha=0.5, filter_strides=(1, 2)))
    for k in ("electric_field", "magnetic_field", "current_density", "charge_density", "positions", "velocities"):
        err = np.abs(out[k] - ref[k]).max() / np.abs(ref[k]).max()
        assert err < 1e-5, (k, err)
    np.testing.assert_allclose(out["initial_velocities"], ref["initial_velocities"], rtol=1e-14)
    np.testing.assert_allclose(out["time_array"], np.linspace(0, 15 * dt, 15))


@pytest.mark.gpu
def test_large_run_without_particle_history_uses_the_binned_engine():
    par = {"domain_parameters": {"number_grid_points": 256, "total_steps": 10, "length": 0.05},
           "species_parameters": {"electrons": {"number_pseudoparticles": 600_000, "random_positions_x": True, "drift_speed_x": 5e7},
                                  "ions": {"number_pseudoparticles": 600_000, "random_positions_x": True}},
           "solver_parameters": {"print_info": False, "particle_history": False}}
    out = Simulation(par).run()
    assert out["positions"] is None and out["electric_field"].shape == (10, 256, 3) and np.isfinite(out["electric_field"]).all()
    rho_sum = out["charge_density"].sum(axis=1) * out["dx"]
    assert np.abs(rho_sum - out["charges"].sum()).max() < 1e-9 * np.abs(out["charges"]).sum()
    with pytest.raises(JicError):
        diagnostics(out)


def test_species_blocks_of_a_carry():
    """Boris_step recovers the library's species table from the carry's (q, m, q/m) arrays (jaxincell_b200/_algorithms.py)."""
    from jaxincell_b200._algorithms import species_blocks
    q = np.array([-2.0, -2.0, 0.0, 3.0, 3.0, 0.0, 0.0]).reshape(-1, 1)
    m = np.array([1.0, 1.0, 1.0, 5.0, 5.0, 7.0, 7.0]).reshape(-1, 1)
    qm = np.array([-2.0, -2.0, 0.0, 0.6, 0.6, 0.0, 0.0]).reshape(-1, 1)
    blocks = species_blocks(q, m, qm)
    assert [(b["count"], b["q"], b["m"], b["qm"]) for b in blocks] == [(3, -2.0, 1.0, -2.0), (2, 3.0, 5.0, 0.6), (2, 0.0, 7.0, 0.0)]
    assert species_blocks(np.zeros(0), np.zeros(0), np.zeros(0))[0]["count"] == 0
    with pytest.raises(JicError):
        species_blocks(np.array([1.0, 2.0]), np.array([1.0, 1.0]), np.array([1.0, 2.0]))   # two charges inside one mass block
    with pytest.raises(JicError):
        species_blocks(np.ones(20), np.arange(20.0), np.ones(20))                          # more blocks than the library takes


def test_section_setters_reclean_and_invalidate_state():
    """Mirrors the reference's tests/test_simulation.py:506-577 (property setters) and :622-660 (current_domain_state): assigning a
    section re-cleans it with the defaults overlaid on the new dictionary, the matching hash changes, cached state follows."""
    par = {"domain_parameters": {"total_steps": 2, "number_grid_points": 4},
           "species_parameters": {"electrons": {"electrons0": {"number_pseudoparticles": 4}}, "ions": {"ions0": {"number_pseudoparticles": 4}}},
           "solver_parameters": {"print_info": False, "rng": "numpy"}}
    sim = Simulation(par)
    hashes = (sim.domain_hash, sim.species_hash, sim.external_field_hash, sim.source_hash, sim.solver_hash)
    assert sim.grid.shape == (4,) and sim.positions.shape == (8, 3) and len(sim.species_index) == 8
    sim.domain_parameters = {"total_steps": 2, "number_grid_points": 5, "length": 0.02}
    assert sim.domain_hash != hashes[0] and sim.grid.shape == (5,) and sim.domain_parameters["length"] == 0.02
    assert sim.domain_parameters["particle_BC_left"] == 0  # defaults overlaid
    sim.species_parameters = {"electrons": {"electrons0": {"number_pseudoparticles": 3}}, "ions": {"ions0": {"number_pseudoparticles": 2}}}
    assert sim.species_hash != hashes[1] and sim.positions.shape == (5, 3) and len(sim.species_index) == 5
    assert sim.species_index[0] == "electrons._electrons0" and sim.species_index[-1] == "ions._ions0"
    sim.external_field_parameters = {"external_electric_field_amplitude": 2.0, "external_electric_field_wavenumber": 1.0}
    assert sim.external_field_hash != hashes[2]
    assert sim.external_electric_field.shape == (5, 3) and sim.external_magnetic_field.dtype == np.float32
    sim.source_parameters = {"source_term_active": 1, "source_species": 0}
    assert sim.source_hash != hashes[3] and sim.source_parameters["injection_speed_x"] == 1e7
    sim.solver_parameters = {"field_solver": 0, "filter_passes": 0, "filter_alpha": 0.25, "print_info": False, "seed": 123, "rng": "numpy"}
    assert sim.solver_hash != hashes[4] and sim.solver_parameters["filter_strides"] == (1, 2, 4)
    assert sim.domain_parameters["number_grid_points"] == 5 and sim.positions.shape == (5, 3)
    state = sim.current_domain_state()
    assert state["dx"] == sim.dx and state["dt"] == sim.dt and tuple(state["box_size"]) == tuple(sim.box_size)
    state["dx"] = -1.0
    assert sim.dx > 0
    with pytest.raises(AssertionError):
        sim.solver_parameters = {"filter_alpha": 2.0}


def test_input_parameters_entry_and_setter_route_like_the_reference():
    """Behaviour observed on the reference's own Simulation (run on tests/refshim; cf. its tests/test_simulation.py:579-620): the
    `input_parameters` entry is routed into the sections; only the reference's differentiable parameters stay exposed; the setter re-routes
    on top of the base sections, into which the NON-differentiable ones of earlier inputs were merged for good."""
    par = {"domain_parameters": {"total_steps": 2, "number_grid_points": 4},
           "species_parameters": {"electrons": {"electrons0": {"number_pseudoparticles": 4}}, "ions": {"ions0": {"number_pseudoparticles": 4}}},
           "solver_parameters": {"print_info": False, "rng": "numpy"},
           "input_parameters": {"length": 0.03, "filter_passes": 2,
                                "electrons": {"electrons0": {"vth_over_c_x": 0.07, "velocity_plus_minus_x": False}}}}
    sim = Simulation(par)
    assert sim.input_parameters == {"electrons": {"electrons0": {"vth_over_c_x": 0.07}}, "length": 0.03}
    e0 = sim.species_parameters["electrons"]["_electrons0"]
    assert (sim.domain_parameters["length"], sim.solver_parameters["filter_passes"], e0["vth_over_c_x"], e0["velocity_plus_minus_x"]) == (0.03, 2, 0.07, False)
    sim.input_parameters = {"length": 0.02, "ions": {"ions0": {"mass_over_proton_mass": 2.0, "number_pseudoparticles": 3}}}
    assert sim.input_parameters == {"ions": {"ions0": {"mass_over_proton_mass": 2.0}}, "length": 0.02}
    assert sim.domain_parameters["length"] == 0.02 and sim.solver_parameters["filter_passes"] == 2       # the non-differentiable one stayed
    assert sim.species_parameters["ions"]["_ions0"]["number_pseudoparticles"] == 3 and sim.positions.shape == (7, 3)
    assert sim.species_parameters["electrons"]["_electrons0"]["vth_over_c_x"] == 0.05                        # the differentiable one did not
    assert sim.species_parameters["electrons"]["_electrons0"]["velocity_plus_minus_x"] is False
    for bad in ({"ion_drift_speed_x": 1.0}, {"nonsense": 1}):
        with pytest.raises(ValueError, match="could not be routed"):
            sim.input_parameters = bad
    Simulation({"nonsense_section": {}})  # top-level keys that are not sections are ignored, as in the reference
