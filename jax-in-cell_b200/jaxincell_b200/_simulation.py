"""Host driver with the reference's public shape: ``Simulation(parameters_or_toml).run()`` -> output dictionary, ``load_parameters``,
``diagnostics``.  It mirrors jaxincell/_simulation.py:36-344 (interface, defaults, output keys) on top of the C ABI; the
per-step arithmetic all happens in libjic_b200.so.  Host code is NumPy: there is no JAX here, so

  * ``run(input_parameters)`` admits what the reference admits there -- its differentiable parameters, same routing, same errors
    (jaxincell/_routing.py:160-226) -- but as plain numbers: no autodiff through the path (out of scope per BASELINE.json);
  * random initial particles are generated on the device by ``jic_sample_particles`` (csrc/jic_sample.cuh): the reference's
    formulas and seed schedule (_state_initialization.py:51-96) on a restatement of ``jax.random``'s Threefry streams
    (``rng="threefry"``, the default; ``threefry_partitionable`` picks jax's bit layout, True = jax >= 0.5).  ``rng="numpy"``
    draws them on the host from ``numpy.random.default_rng`` instead (same distributions, different streams).  Explicit
    ``initial_positions`` / ``initial_velocities`` per species override either (_parameters/_species_definitions.py:57-58);
  * ``field_solver = 1`` (the per-step Gauss correction of _algorithms.py:69-78) and ``time_evolution_algorithm = 1`` (the implicit
    Crank-Nicolson stepper, _algorithms.py:100-241) both run in the library (k_gauss; csrc/jic_cn.cuh).

Extra, optional ``solver_parameters`` understood here only: ``engine`` ("auto" | "indexed" | "binned"), ``particle_history``
(default True, like the reference; False drops the (T,N,3) histories and lets large runs use the binned engine), ``dtype``
("float64" | "float32"), ``device_resident`` ("auto" | True | False) and ``kinetic_energy_history`` (default False).

Large runs (``particle_history=False``; a million particles or more, or several ranks) take the DEVICE-RESIDENT path: like the
reference, which builds its particles inside the jit on the device (jaxincell/_simulation.py:169-190), the particles are sampled
on the GPU and handed to the library there -- no host copy of the (N,3) arrays exists unless the output dictionary is asked to
hold them (N <= ``PARTICLE_ARRAYS_LIMIT``).  Under ``torchrun`` (an initialised ``torch.distributed`` NCCL group) every rank
samples only its own index slice of every species (Threefry is counter based), pushes it, and the library reduces the grid
over NVLink every step: fields and energies in the output are global, per-particle arrays (when present) are the rank's slice.
``kinetic_energy_history=True`` adds a per-species kinetic-energy history reduced on the device (``kinetic_energy_species``,
(T, n_species)), which ``diagnostics`` uses when the velocity history does not exist.
"""
from __future__ import annotations

import copy
import os

import numpy as np

from ._lib import JicError

try:  # jaxincell/_simulation.py:27-30
    import tomllib
except ModuleNotFoundError:  # pragma: no cover
    import pip._vendor.tomli as tomllib

# jaxincell/_constants.py:1-7 (the rounded values are part of the results)
epsilon_0 = 8.85418782e-12
mu_0 = 1.25663706e-6
speed_of_light = 2.99792458e8
elementary_charge = 1.60217663e-19
mass_electron = 9.10938371e-31
mass_proton = 1.67262193e-27

AXES = ("x", "y", "z")
SECTIONS = ("domain_parameters", "species_parameters", "external_field_parameters", "source_parameters", "solver_parameters")

# defaults: _parameters/_domain_parameters.py:12-25, _solver_parameters.py:10-22, _external_field_parameters.py:10-17
DOMAIN_DEFAULTS = dict(total_steps=350, timestep_over_spatialstep_times_c=1.0, number_grid_points=50, number_grid_points_y=0,
                       number_grid_points_z=0, length=1e-2, length_y=0, length_z=0, particle_BC_left=0, particle_BC_right=0,
                       field_BC_left=0, field_BC_right=0)
SOLVER_DEFAULTS = dict(print_info=True, field_solver=0, relativistic=False, time_evolution_algorithm=0,
                       max_number_of_Picard_iterations_implicit_CN=20, number_of_particle_substeps_implicit_CN=2,
                       tolerance_Picard_iterations_implicit_CN=1e-6, filter_passes=5, filter_alpha=0.5, filter_strides=(1, 2, 4),
                       seed=1701, engine="auto", particle_history=True, dtype="float64", rng="threefry", threefry_partitionable=True,
                       device_resident="auto", kinetic_energy_history=False)
PARTICLE_ARRAYS_LIMIT = 20_000_000  # device-resident runs above this hold None under the per-particle output keys
# _parameters/_source_parameters.py:10-23 -- carried into the output dictionary; no code of the step reads them
SOURCE_DEFAULTS = dict(source_term_active=0, source_species=1, how_often_source_should_produce_quasiparticles=20,
                       source_particles_per_second=1e16, location_of_source=0, width_of_source=1, injection_speed_x=1e7,
                       injection_speed_y=0, injection_speed_z=0)
EXTERNAL_DEFAULTS = dict(external_electric_field_amplitude=0.0, external_electric_field_wavenumber=0.0,
                         external_magnetic_field_amplitude=0.0, external_magnetic_field_wavenumber=0.0,
                         external_electric_field_function=None, external_magnetic_field_function=None)


def _species_defaults(kind, first):
    """_parameters/_species_definitions.py:33-101: the first species of each type gets the two-stream-like 'initial' defaults."""
    d = dict(number_pseudoparticles=500, grid_points_per_Debye_length=2, weight=0, seed_position_override=False, seed_position=None,
             initial_positions=None, initial_velocities=None)
    for a in AXES:
        d.update({f"perturbation_amplitude_{a}": 0.0, f"perturbation_wavenumber_{a}": 0, f"random_positions_{a}": a != "x",
                  f"vth_over_c_{a}": 0, f"drift_speed_{a}": 0, f"velocity_plus_minus_{a}": False})
    if kind == "electrons":
        d["charge_over_elementary_charge"] = -1
        if first:
            d.update(perturbation_amplitude_x=1e-7, perturbation_wavenumber_x=8, vth_over_c_x=0.05, drift_speed_x=1e8, velocity_plus_minus_x=True)
    else:
        d.update(charge_over_elementary_charge=1, mass_over_proton_mass=1)
        for a in AXES:
            d[f"ion_temperature_over_electron_temperature_{a}"] = 1
        if first:
            d.update(perturbation_amplitude_x=1e-7, vth_over_c_x="_electrons0", vth_over_c_y="_electrons0", vth_over_c_z="_electrons0")
    return d


def load_parameters(input_file):
    """jaxincell/_simulation.py:36-41"""
    with open(input_file, "rb") as f:
        return tomllib.load(f)


def _normalize_species_input(species_parameters):
    """{kind: {user_label: values}} whatever the input shape (_species_parameters.py:52-64): empty -> one default species, flat values ->
    the default label."""
    norm = {}
    for kind in ("electrons", "ions"):
        values = dict(species_parameters.get(kind, {}) or {})
        if not values:
            values = {f"{kind}0": {}}
        elif not any(isinstance(v, dict) for v in values.values()):
            values = {f"{kind}0": values}
        norm[kind] = {label: dict(v) for label, v in values.items()}
    return norm


def _clean_species(species_parameters):
    """Canonical labels `_electrons<i>` / `_ions<i>`, defaults, cross references (_species_parameters.py:52-200)."""
    out = {}
    for kind, values in _normalize_species_input(species_parameters).items():
        out[kind] = {}
        for i, (label, v) in enumerate(values.items()):
            out[kind][f"_{kind}{i}"] = {**_species_defaults(kind, i == 0), **v, "user_label": label}
    labels = {k: {sp["user_label"]: canon for canon, sp in out[k].items()} for k in out}

    def find(ref):
        for k in ("ions", "electrons"):
            if ref in out[k]:
                return k, ref
            if ref in labels[k]:
                return k, labels[k][ref]
        raise ValueError(f"Cross referenced species value did not reference another species: {ref!r}")

    for kind in out:
        for canon, sp in out[kind].items():
            for key, val in list(sp.items()):
                if key == "user_label" or not isinstance(val, str):
                    continue
                rk, rl = find(val)
                ref = out[rk][rl]
                if isinstance(ref[key], str):
                    raise AssertionError("Cross referenced species values cannot reference another reference.")
                if key.startswith("vth_over_c_") and kind == "ions" and rk == "electrons":
                    a = key[-1]
                    sp[key] = float(np.sqrt(sp[f"ion_temperature_over_electron_temperature_{a}"]) * ref[key]
                                    * np.sqrt(mass_electron / (sp["mass_over_proton_mass"] * mass_proton)))
                elif key.startswith("vth_over_c_") and kind == "electrons" and rk == "ions":
                    a = key[-1]
                    sp[key] = float(np.sqrt(1 / ref[f"ion_temperature_over_electron_temperature_{a}"]) * ref[key]
                                    * np.sqrt(ref["mass_over_proton_mass"] * mass_proton / mass_electron))
                else:
                    sp[key] = ref[key]
    # validation after the references are resolved, as in the reference (_species_parameters.py:66-100,213-221)
    for kind in out:
        for canon, sp in out[kind].items():
            if not (type(sp["number_pseudoparticles"]) == int and sp["number_pseudoparticles"] > 0):
                raise AssertionError(f"Number of pseudoparticles for {canon} must be a positive integer. Got {sp['number_pseudoparticles']}.")
            if not sp["grid_points_per_Debye_length"] > 0:
                raise AssertionError(f"Grid points per Debye length must be positive. Got {sp['grid_points_per_Debye_length']}.")
            if not sp["weight"] >= 0:
                raise AssertionError(f"Weight must be non-negative. Got {sp['weight']}.")
            if kind == "ions":
                if not sp["mass_over_proton_mass"] > 0:
                    raise AssertionError(f"Mass over proton mass must be positive. Got {sp['mass_over_proton_mass']}.")
                for a in AXES:
                    if not sp[f"ion_temperature_over_electron_temperature_{a}"] >= 0:
                        raise AssertionError(f"Ion temperature over electron temperature {a} must be positive.")
            for key in [f"{k}_{a}" for k in ("random_positions", "velocity_plus_minus") for a in AXES] + ["seed_position_override"]:
                if type(sp[key]) != bool:
                    raise AssertionError(f"{key} must be a boolean. Got {sp[key]}.")
            if sp["seed_position_override"] and not (type(sp["seed_position"]) == int and sp["seed_position"] > 0):
                raise AssertionError(f"Seed position must be a positive integer. Got {sp['seed_position']}.")
    return out


# what `run(input_parameters)` may override (the reference's DIFFERENTIABLE_* lists: _domain_parameters.py:27-32, _solver_parameters.py:24-26,
# _species_definitions.py:113-146); everything else is fixed at construction, as in the reference
RUNTIME_FLAT_KEYS = {"timestep_over_spatialstep_times_c": "domain_parameters", "length": "domain_parameters", "length_y": "domain_parameters",
                     "length_z": "domain_parameters", "filter_alpha": "solver_parameters"}
_RUNTIME_SPECIES_COMMON = ("grid_points_per_Debye_length", "weight", "charge_over_elementary_charge",
                           *(f"{k}_{a}" for k in ("perturbation_amplitude", "perturbation_wavenumber", "vth_over_c", "drift_speed") for a in AXES),
                           "initial_positions", "initial_velocities")
RUNTIME_SPECIES_KEYS = {"electrons": _RUNTIME_SPECIES_COMMON,
                        "ions": _RUNTIME_SPECIES_COMMON + ("mass_over_proton_mass", *(f"ion_temperature_over_electron_temperature_{a}" for a in AXES))}


def _seed_pair(seed, kind, rng_index, extra):
    """_state_initialization.py:87-96"""
    if rng_index == 0:
        return (seed, seed + 3) if kind == "electrons" else (seed, seed + 6)
    s = seed + 12 + extra * 6
    return s, s


class Simulation:
    """Same call shape as jaxincell.Simulation (jaxincell/_simulation.py:43-121)."""

    def __init__(self, parameters=None):
        if parameters is None:
            parameters = {}
        if not isinstance(parameters, dict):
            parameters = load_parameters(parameters)
        self._ingest(parameters)

    def _ingest(self, parameters):
        """Constructor and `input_parameters` setter (jaxincell/_simulation.py:346-417, _routing.py:77-132): the `input_parameters` entry is
        routed into the sections; the reference's differentiable parameters among them stay visible as `sim.input_parameters` and are an
        overlay on the base sections, everything else becomes part of the base sections.  A flat key no section knows is an error; top
        level keys that are not sections are ignored, as in the reference."""
        parameters = copy.deepcopy(parameters)
        given = dict(parameters.pop("input_parameters", {}) or {})
        base = {name: dict(parameters.get(name, {}) or {}) for name in SECTIONS}
        base["species_parameters"] = _normalize_species_input(base["species_parameters"])
        overlay = copy.deepcopy(base)
        exposed, unrouted = {}, []
        flat_defaults = (("domain_parameters", DOMAIN_DEFAULTS), ("solver_parameters", SOLVER_DEFAULTS),
                         ("external_field_parameters", EXTERNAL_DEFAULTS), ("source_parameters", SOURCE_DEFAULTS))
        for key, val in given.items():
            if key in ("electrons", "ions"):
                labels = list(base["species_parameters"][key])
                canonical = {f"_{key}{i}": lab for i, lab in enumerate(labels)}
                groups = val.items() if any(isinstance(v, dict) for v in val.values()) else [(None, val)]
                for label, over in groups:
                    if label is None:
                        targets = labels
                    elif label in canonical or label in labels:
                        targets = [canonical.get(label, label)]
                    else:  # a new species introduced through input_parameters
                        base["species_parameters"][key][label] = {}
                        overlay["species_parameters"][key][label] = {}
                        targets = [label]
                    for t in targets:
                        for k, v in over.items():
                            overlay["species_parameters"][key][t][k] = v
                            if k in RUNTIME_SPECIES_KEYS[key]:
                                exposed.setdefault(key, {}).setdefault(t if label is None else label, {})[k] = v
                            else:
                                base["species_parameters"][key][t][k] = v
                continue
            section = next((name for name, defaults in flat_defaults if key in defaults), None)
            if section is None:
                unrouted.append(key)
                continue
            overlay[section][key] = val
            if key in RUNTIME_FLAT_KEYS:
                exposed[key] = val
            else:
                base[section][key] = val
        if unrouted:
            raise ValueError("Initial input_parameters included parameter(s) that could not be routed. Unrouted parameter(s): " + ", ".join(unrouted))
        self._base_sections = base
        self._input_parameters = exposed
        for name in SECTIONS:  # the section setters clean and overlay the defaults, as the reference's do
            self._set_section(name, overlay[name], remember=False)

    @property
    def input_parameters(self):
        """The differentiable parameters that came in through `input_parameters` (_simulation.py:548-550)."""
        return copy.deepcopy(self._input_parameters)

    @input_parameters.setter
    def input_parameters(self, new_input_parameters):
        """Re-route on top of the base sections (_simulation.py:552-558)."""
        parameters = copy.deepcopy(self._base_sections)
        parameters["input_parameters"] = new_input_parameters
        self._ingest(parameters)

    # ---- parameter sections as properties: assigning one re-cleans it (defaults overlaid on the NEW dict only) and thereby invalidates
    #      the cached state, like set_parameter_section of the reference (jaxincell/_simulation.py:494-556)
    def _set_section(self, name, new, remember=True):
        new = copy.deepcopy(dict(new or {}))
        if remember:  # an assignment from outside replaces the base section (set_parameter_section, _simulation.py:494-506)
            self._base_sections[name] = _normalize_species_input(new) if name == "species_parameters" else copy.deepcopy(new)
        if name == "domain_parameters":
            self._domain_parameters = self._clean_domain({**DOMAIN_DEFAULTS, **new})
        elif name == "solver_parameters":
            self._solver_parameters = self._clean_solver({**SOLVER_DEFAULTS, **new})
        elif name == "external_field_parameters":
            self._external_field_parameters = {**EXTERNAL_DEFAULTS, **new}
        elif name == "source_parameters":
            self._source_parameters = {**SOURCE_DEFAULTS, **new}
        else:
            self._species_input = _normalize_species_input(new)  # unresolved: runtime overrides re-resolve the cross references
            self._species_parameters = _clean_species(self._species_input)

    domain_parameters = property(lambda self: self._domain_parameters, lambda self, v: self._set_section("domain_parameters", v))
    species_parameters = property(lambda self: self._species_parameters, lambda self, v: self._set_section("species_parameters", v))
    external_field_parameters = property(lambda self: self._external_field_parameters, lambda self, v: self._set_section("external_field_parameters", v))
    source_parameters = property(lambda self: self._source_parameters, lambda self, v: self._set_section("source_parameters", v))
    solver_parameters = property(lambda self: self._solver_parameters, lambda self, v: self._set_section("solver_parameters", v))

    @staticmethod
    def _hash(section):
        """build_parameter_hash of the reference (_parameters/_utils.py:15-20): changes whenever a value of the section changes."""
        return "".join(str(k) + str(v) for k, v in section.items())

    domain_hash = property(lambda self: self._hash(self._domain_parameters))
    species_hash = property(lambda self: str(tuple((kind, tuple(label + self._hash(sp) for label, sp in group.items()))
                                                     for kind, group in self._species_parameters.items())))
    external_field_hash = property(lambda self: self._hash(self._external_field_parameters))
    source_hash = property(lambda self: self._hash(self._source_parameters))
    solver_hash = property(lambda self: self._hash(self._solver_parameters))

    def current_domain_state(self):
        """jaxincell/_simulation.py:430-436"""
        return self.build_domain_state(self.domain_parameters)

    # ---- state attributes of the reference's Simulation object (jaxincell/_simulation.py:438-492): dx, dt, grid, box_size,
    #      positions, velocities, weights, charges, masses, ..., fields, external_*_field -- computed on first use, cached per parameter set
    _DOMAIN_ATTRS = ("dx", "dt", "grid", "box_size")
    _PARTICLE_ATTRS = ("positions", "velocities", "weights", "charges", "masses", "charge_to_mass_ratios", "species_integer_index",
                       "charge_integer_lookup", "mass_integer_lookup", "charge_mass_integer_lookup", "vth_electrons", "vth_electrons_over_c",
                       "charge_electrons")
    _PARTICLE_ATTRS += ("species_index",)
    _FIELD_ATTRS = ("fields",)
    _EXTERNAL_ATTRS = {"external_electric_field": ("external_electric_field", "E"), "external_magnetic_field": ("external_magnetic_field", "B")}

    def __getattr__(self, name):  # only reached for names that are not regular attributes
        if name in Simulation._DOMAIN_ATTRS:
            return self.build_domain_state(self.domain_parameters)[name]
        if name in Simulation._PARTICLE_ATTRS:
            return self._state()["particles"][name]
        if name in Simulation._FIELD_ATTRS:
            return self._state(fields=True)[name]
        if name in Simulation._EXTERNAL_ATTRS:  # float32 (G,3), zeros unless the section carries {"E": array} / {"B": array} (_state_initialization.py:380-392)
            key, comp = Simulation._EXTERNAL_ATTRS[name]
            given = self.external_field_parameters.get(key)
            G = int(self.domain_parameters["number_grid_points"])
            return np.asarray(given[comp], np.float32) if isinstance(given, dict) and comp in given else np.zeros((G, 3), np.float32)
        raise AttributeError(name)

    def _state(self, fields=False):
        """Particle state (host arrays; random draws come from the device unless rng='numpy') and, on request, the initial fields:
        E_x from Gauss's law on the filtered initial charge (_state_initialization.py:365-398), computed by the library."""
        key = repr((self.domain_parameters, self.solver_parameters, self.species_parameters, sorted(self.external_field_parameters)))
        cache = self.__dict__.get("_state_cache")
        if cache is None or cache["key"] != key:
            dom, solver = self.domain_parameters, self.solver_parameters
            st = self.build_domain_state(dom)
            cache = self.__dict__["_state_cache"] = dict(key=key, domain=st, particles=self.initialize_particle_state(self.species_parameters, dom, solver, st))
        if fields and "fields" not in cache:
            from ._engine import HotPath
            dom, solver, st, ps = self.domain_parameters, self.solver_parameters, cache["domain"], cache["particles"]
            G = int(dom["number_grid_points"])
            hp = HotPath(species=ps["species_table"], length=st["box_size"][0], length_y=st["box_size"][1], length_z=st["box_size"][2], G=G,
                         dt=st["dt"], pbl=dom["particle_BC_left"], pbr=dom["particle_BC_right"], fbl=dom["field_BC_left"], fbr=dom["field_BC_right"],
                         filter_passes=solver["filter_passes"], filter_alpha=solver["filter_alpha"], filter_strides=solver["filter_strides"],
                         engine="indexed")
            try:
                hp.set_external_fields(None, None)
                hp.initialize(ps["positions"], ps["velocities"])
                E0, B0, _ = hp.initial(velocities=False)
                cache["fields"] = (E0.cpu().numpy(), B0.cpu().numpy())
            finally:
                hp.close()
        return cache

    # ---- cleaning -------------------------------------------------------------------------------------------------
    @staticmethod
    def _clean_domain(d):
        assert type(d["total_steps"]) == int and d["total_steps"] > 0, "Total number of time steps must be an integer."
        assert d["length"] > 0, "Length of the simulation box must be positive."
        for k in ("particle_BC_left", "particle_BC_right", "field_BC_left", "field_BC_right"):
            assert d[k] in (0, 1, 2), f"Invalid boundary condition {k}. Must be 0 (periodic), 1 (reflecting), or 2 (absorbing)."
        return d

    @staticmethod
    def _clean_solver(s):
        s["filter_strides"] = tuple(s["filter_strides"])
        assert s["field_solver"] in (0, 1), "Invalid field solver."
        assert s["time_evolution_algorithm"] in (0, 1), "Invalid time evolution algorithm."
        assert s.get("rng", "threefry") in ("threefry", "numpy"), "rng must be 'threefry' (device, jax.random streams) or 'numpy'."
        assert type(s["filter_passes"]) == int and s["filter_passes"] >= 0, "Number of passes of the digital filter must be a non-negative integer."
        assert 0 < s["filter_alpha"] < 1, "Filter strength must be a float between 0 and 1."
        assert all(type(v) == int and v > 0 for v in s["filter_strides"]), "Filter strides must be a tuple of positive integers."
        return s

    def clean_runtime_input_parameters(self, input_parameters=None):
        """jaxincell/_routing.py:160-226: only the reference's differentiable parameters may change between runs of one Simulation;
        species overrides are nested {electrons|ions: {label: {...}}} (canonical or user label; without labels: every species of the
        type).  Same errors as the reference: TypeError for a non-dict, ValueError naming every offending path."""
        if input_parameters is None:
            return {}
        if not isinstance(input_parameters, dict):
            raise TypeError("Runtime input_parameters must be a dictionary.")
        invalid = []
        for key, value in input_parameters.items():
            if key in ("electrons", "ions"):
                if not isinstance(value, dict):
                    raise TypeError(f"Runtime input_parameters[{key!r}] must be a dictionary.")
                groups = value.items() if any(isinstance(v, dict) for v in value.values()) else [(None, value)]
                labels = {lab for canon, sp in self.species_parameters[key].items() for lab in (canon, sp["user_label"])}
                for label, over in groups:
                    if not isinstance(over, dict):
                        raise TypeError(f"Runtime input_parameters for {key} must be species dictionaries or flat values, not a mix.")
                    for k in over:
                        if k not in RUNTIME_SPECIES_KEYS[key]:
                            invalid.append(f"{key}.{k}" if label is None else f"{key}.{label}.{k}")
                    if label is not None and label not in labels:
                        raise ValueError(f"Could not find {key} species {label!r}.")
            elif key not in RUNTIME_FLAT_KEYS:
                invalid.append(key)
        if invalid:
            raise ValueError("Runtime input_parameters can only contain differentiable parameters. Invalid parameter(s): " + ", ".join(invalid))
        return input_parameters

    def _sections(self, input_parameters):
        """Overrides: a key is routed to whichever section declares it; species overrides are nested {electrons|ions: {label: {...}}}."""
        sec = {k: copy.deepcopy(getattr(self, k)) for k in SECTIONS}
        raw_species = None
        for key, val in (input_parameters or {}).items():  # (what came in at construction is part of the sections already)
            if key in ("electrons", "ions"):
                # start from the UNRESOLVED input so that cross references ("_electrons0") follow the overridden values, as in the
                # reference, which re-resolves them inside _simulation (_simulation.py:163, _species_parameters.py:129-166)
                raw_species = raw_species or copy.deepcopy(self._species_input)
                user_labels = list(raw_species[key])
                canonical = {f"_{key}{i}": lab for i, lab in enumerate(user_labels)}
                groups = val.items() if any(isinstance(v, dict) for v in val.values()) else [(None, val)]
                for label, over in groups:
                    if label is None:
                        targets = user_labels  # no label: every species of the type (_routing.py:185-186)
                    elif label in canonical:
                        targets = [canonical[label]]
                    elif label in raw_species[key]:
                        targets = [label]
                    else:
                        raise ValueError(f"Could not find {key} species {label!r}.")
                    for t in targets:
                        raw_species[key][t].update(over)
                continue
            for name in ("domain_parameters", "solver_parameters", "external_field_parameters", "source_parameters"):
                if key in sec[name] or name == "source_parameters":
                    sec[name][key] = val
                    break
        if raw_species is not None:
            sec["species_parameters"] = _clean_species(raw_species)
        sec["domain_parameters"] = self._clean_domain(sec["domain_parameters"])
        sec["solver_parameters"] = self._clean_solver(sec["solver_parameters"])
        return sec

    # ---- state initialisation (NumPy) -----------------------------------------------------------------------------
    @staticmethod
    def build_domain_state(dom):
        """_state_initialization.py:27-49"""
        length = float(dom["length"])
        G = int(dom["number_grid_points"])
        dx = length / G
        return dict(box_size=(length, float(dom["length_y"]) or length, float(dom["length_z"]) or length), dx=dx,
                    dt=dom["timestep_over_spatialstep_times_c"] * dx / speed_of_light,
                    grid=np.linspace(-length / 2 + dx / 2, length / 2 - dx / 2, G))

    @staticmethod
    def initialize_particle_state(species_parameters, dom, solver, state, materialize=True):
        """_state_initialization.py:51-286.  Random draws: device Threefry (jic_sample_particles) or numpy.random.default_rng.
        materialize=False (device-resident runs): only the per-species tables; no per-particle host array is built."""
        box, G = state["box_size"], int(dom["number_grid_points"])
        threefry = solver.get("rng", "threefry") == "threefry"
        if not materialize:
            return Simulation._species_level_state(species_parameters, dom, solver, state)
        sampling, given_any = [], []
        pos, vel, wts, sidx, table, names = [], [], [], [], [], []
        charge_l, mass_l, qm_l = [], [], []
        ref = None
        extra = 0
        for kind in ("electrons", "ions"):
            for i, (canon, sp) in enumerate(species_parameters[kind].items()):
                sp_seed, sv_seed = _seed_pair(int(solver["seed"]), kind, i, extra)
                if i != 0:
                    extra += 1
                if sp["seed_position_override"]:
                    sp_seed = sp["seed_position"]
                n = sp["number_pseudoparticles"]
                x = np.empty((n, 3)); v = np.empty((n, 3))
                sampling.append(dict(count=n, seed_position=sp_seed, seed_velocity=sv_seed,
                                     **{k: [sp[f"{k}_{ax}"] for ax in AXES] for k in
                                        ("random_positions", "velocity_plus_minus", "perturbation_amplitude", "perturbation_wavenumber",
                                         "vth_over_c", "drift_speed")}))
                for a, ax in enumerate(AXES if not threefry else ()):
                    if sp[f"random_positions_{ax}"]:
                        xa = np.random.default_rng(sp_seed + a + 1).uniform(-box[a] / 2, box[a] / 2, n)
                    else:
                        xa = np.linspace(-box[a] / 2, box[a] / 2, n)
                    xa = xa + sp[f"perturbation_amplitude_{ax}"] * np.sin(sp[f"perturbation_wavenumber_{ax}"] * 2 * np.pi / box[a] * xa)
                    va = sp[f"vth_over_c_{ax}"] * speed_of_light / np.sqrt(2) * np.random.default_rng(sv_seed + a + 4).standard_normal(n)
                    va = va + sp[f"drift_speed_{ax}"]
                    if sp[f"velocity_plus_minus_{ax}"]:
                        va = va * (-1.0) ** np.arange(n)
                    x[:, a], v[:, a] = xa, va
                for key, arr in (("initial_positions", x), ("initial_velocities", v)):
                    if sp[key] is not None:
                        given = np.asarray(sp[key], dtype=float)
                        assert given.shape == (n, 3), f"{key} for {kind}{i} must have shape {(n, 3)}. Got {given.shape}."
                        arr[...] = given
                        given_any.append((len(pos), key))
                mass = mass_electron if kind == "electrons" else sp["mass_over_proton_mass"] * mass_proton
                charge = sp["charge_over_elementary_charge"] * elementary_charge
                if kind == "electrons" and i == 0:
                    vths = [sp[f"vth_over_c_{ax}"] for ax in AXES]
                    ref = dict(vth_electrons=max(vths) * speed_of_light, vth_electrons_over_c=max(vths), charge_electrons=charge)
                if ref is None:
                    raise ValueError("Electron reference species must be initialized before ions.")
                debye_length_per_dx = 1 / sp["grid_points_per_Debye_length"]  # same operation order as the reference: the weight is bit-equal
                w = (epsilon_0 * mass_electron * speed_of_light ** 2 / ref["charge_electrons"] ** 2 * G ** 2 / box[0] / (2 * n)
                     * ref["vth_electrons_over_c"] ** 2 / debye_length_per_dx ** 2)
                w = w if sp["weight"] == 0 else float(sp["weight"])
                pos.append(x); vel.append(v); wts.append(np.full((n, 1), w)); sidx.append(np.full(n, len(table), dtype=np.int32))
                names.extend([f"{kind}.{canon}"] * n)
                table.append(dict(count=n, q=charge * w, m=mass * w, qm=charge / mass))
                charge_l.append(charge); mass_l.append(mass); qm_l.append(charge / mass)
        if threefry:
            # one device call for all species; explicit initial_positions / initial_velocities then replace their blocks
            from ._engine import sample_particles
            dx0, dv0 = sample_particles(sampling, box, threefry_partitionable=bool(solver.get("threefry_partitionable", True)))
            hx, hv = dx0.cpu().numpy(), dv0.cpu().numpy()
            o = 0
            for k, (x, v) in enumerate(zip(pos, vel)):
                n = len(x)
                if (k, "initial_positions") not in given_any:
                    x[...] = hx[o:o + n]
                if (k, "initial_velocities") not in given_any:
                    v[...] = hv[o:o + n]
                o += n
        positions, velocities = np.concatenate(pos), np.concatenate(vel)
        weights, species_integer_index = np.concatenate(wts), np.concatenate(sidx)
        lim = 0.99 * speed_of_light
        velocities = np.where(np.abs(velocities) >= lim, np.sign(velocities) * lim, velocities)
        cl, ml, ql = np.array(charge_l), np.array(mass_l), np.array(qm_l)
        return dict(positions=positions, velocities=velocities, weights=weights, species_integer_index=species_integer_index,
                    charge_integer_lookup=cl, mass_integer_lookup=ml, charge_mass_integer_lookup=ql,
                    charges=cl[species_integer_index].reshape(-1, 1) * weights, masses=ml[species_integer_index].reshape(-1, 1) * weights,
                    charge_to_mass_ratios=ql[species_integer_index].reshape(-1, 1), species_table=table, sampling=sampling,
                    species_index=names, **ref)

    @staticmethod
    def _species_level_state(species_parameters, dom, solver, state):
        """The per-species part of initialize_particle_state (same seeds, weights, charges: same code path of the reference,
        _state_initialization.py:87-96,172-185,242-261) without any per-particle array."""
        box, G = state["box_size"], int(dom["number_grid_points"])
        sampling, table, overrides, weights = [], [], {}, []
        charge_l, mass_l, qm_l = [], [], []
        ref, extra = None, 0
        for kind in ("electrons", "ions"):
            for i, (canon, sp) in enumerate(species_parameters[kind].items()):
                sp_seed, sv_seed = _seed_pair(int(solver["seed"]), kind, i, extra)
                if i != 0:
                    extra += 1
                if sp["seed_position_override"]:
                    sp_seed = sp["seed_position"]
                n = sp["number_pseudoparticles"]
                sampling.append(dict(count=n, seed_position=sp_seed, seed_velocity=sv_seed,
                                     **{k: [sp[f"{k}_{ax}"] for ax in AXES] for k in
                                        ("random_positions", "velocity_plus_minus", "perturbation_amplitude", "perturbation_wavenumber",
                                         "vth_over_c", "drift_speed")}))
                for key in ("initial_positions", "initial_velocities"):
                    if sp[key] is not None:
                        given = np.asarray(sp[key], dtype=float)
                        assert given.shape == (n, 3), f"{key} for {kind}{i} must have shape {(n, 3)}. Got {given.shape}."
                        overrides.setdefault(len(table), {})[key] = given
                mass = mass_electron if kind == "electrons" else sp["mass_over_proton_mass"] * mass_proton
                charge = sp["charge_over_elementary_charge"] * elementary_charge
                if kind == "electrons" and i == 0:
                    vths = [sp[f"vth_over_c_{ax}"] for ax in AXES]
                    ref = dict(vth_electrons=max(vths) * speed_of_light, vth_electrons_over_c=max(vths), charge_electrons=charge)
                if ref is None:
                    raise ValueError("Electron reference species must be initialized before ions.")
                debye_length_per_dx = 1 / sp["grid_points_per_Debye_length"]
                w = (epsilon_0 * mass_electron * speed_of_light ** 2 / ref["charge_electrons"] ** 2 * G ** 2 / box[0] / (2 * n)
                     * ref["vth_electrons_over_c"] ** 2 / debye_length_per_dx ** 2)
                w = w if sp["weight"] == 0 else float(sp["weight"])
                weights.append(w)
                table.append(dict(count=n, q=charge * w, m=mass * w, qm=charge / mass, name=f"{kind}.{canon}"))
                charge_l.append(charge); mass_l.append(mass); qm_l.append(charge / mass)
        return dict(species_table=table, sampling=sampling, overrides=overrides, species_weights=weights,
                    charge_integer_lookup=np.array(charge_l), mass_integer_lookup=np.array(mass_l),
                    charge_mass_integer_lookup=np.array(qm_l), **ref)

    @staticmethod
    def _distributed():
        """(rank, world) of an initialised torch.distributed group, else (0, 1)."""
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                return dist.get_rank(), dist.get_world_size()
        except Exception:  # noqa: BLE001
            pass
        return 0, 1

    def _simulation_device(self, sec, state, rank, world):
        """The run with the particles resident on the device from the first random bit to the last step (module docstring)."""
        import torch
        from ._engine import HotPath, sample_particles
        from ._parallel import shard_counts, shard_species
        dom, solver, ext = sec["domain_parameters"], sec["solver_parameters"], sec["external_field_parameters"]
        ps = self.initialize_particle_state(sec["species_parameters"], dom, solver, state, materialize=False)
        G, T = int(dom["number_grid_points"]), int(dom["total_steps"])
        box = state["box_size"]
        tdt = torch.float64 if str(solver["dtype"]) in ("float64", "f64") else torch.float32
        if not torch.cuda.is_available():
            raise JicError("no CUDA device: jaxincell_b200 has no CPU path")
        device = torch.device("cuda", torch.cuda.current_device())
        eE = ext.get("external_electric_field"); eB = ext.get("external_magnetic_field")
        ext_E = np.asarray(eE["E"], np.float32) if isinstance(eE, dict) and "E" in eE else np.zeros((G, 3), np.float32)
        ext_B = np.asarray(eB["B"], np.float32) if isinstance(eB, dict) and "B" in eB else np.zeros((G, 3), np.float32)
        if ext_E.shape != (G, 3) or ext_B.shape != (G, 3):
            raise JicError(f"external fields must have shape ({G}, 3), got {ext_E.shape} and {ext_B.shape}")
        x0, v0 = sample_particles(ps["sampling"], box, dtype=tdt, device=device,
                                  threefry_partitionable=bool(solver.get("threefry_partitionable", True)), rank=rank, world=world)
        o = 0
        for k, spl in enumerate(ps["sampling"]):  # explicit per-species arrays replace this rank's slice of their block
            counts = shard_counts(spl["count"], world)
            lo, n = sum(counts[:rank]), counts[rank]
            for key, arr in ps["overrides"].get(k, {}).items():
                t = torch.as_tensor(arr[lo:lo + n]).to(device=device, dtype=tdt)
                if key == "initial_velocities":
                    lim = 0.99 * speed_of_light
                    t = torch.where(t.abs() >= lim, torch.sign(t) * lim, t)
                (x0 if key == "initial_positions" else v0)[o:o + n] = t
            o += n
        engine = solver["engine"] if solver["engine"] in ("indexed", "binned") else "binned"
        local_table = shard_species(ps["species_table"], rank, world)
        hp = HotPath(species=local_table, dtype=tdt, device=device, engine=engine,
                     length=box[0], length_y=box[1], length_z=box[2], G=G, dt=state["dt"],
                     pbl=dom["particle_BC_left"], pbr=dom["particle_BC_right"], fbl=dom["field_BC_left"], fbr=dom["field_BC_right"],
                     filter_passes=solver["filter_passes"], filter_alpha=solver["filter_alpha"], filter_strides=solver["filter_strides"],
                     relativistic=bool(solver["relativistic"]), field_solver=int(solver["field_solver"]),
                     time_evolution_algorithm=int(solver["time_evolution_algorithm"]),
                     cn_substeps=int(solver["number_of_particle_substeps_implicit_CN"]),
                     cn_max_iterations=int(solver["max_number_of_Picard_iterations_implicit_CN"]),
                     cn_tolerance=float(solver["tolerance_Picard_iterations_implicit_CN"]))
        try:
            if world > 1:
                hp.comm_init_from_torch()
            hp.set_external_fields(ext_E, ext_B)
            hp.initialize(x0, v0)
            keep = hp.N <= PARTICLE_ARRAYS_LIMIT
            hx0 = x0.cpu().numpy() if keep else None
            hv0 = v0.cpu().numpy() if keep else None
            del x0, v0
            E0, B0, _ = hp.initial(velocities=False)
            ke = bool(solver.get("kinetic_energy_history", False))
            out = hp.run(T, kinetic=ke)
            hp.check_status()
            res = {k: out[k].cpu().numpy() for k in ("electric_field", "magnetic_field", "current_density", "charge_density")}
            res["fields"] = (E0.cpu().numpy(), B0.cpu().numpy())
            if ke:
                kin = out["kinetic_energy"]
                if world > 1:
                    import torch.distributed as dist
                    dist.all_reduce(kin)
                res["kinetic_energy_species"] = kin.cpu().numpy()
        finally:
            hp.close()
        # per-particle output arrays: this rank's slice, or None above PARTICLE_ARRAYS_LIMIT
        if keep:
            counts = [s_["count"] for s_ in local_table]
            sidx = np.repeat(np.arange(len(counts), dtype=np.int32), counts)
            w = np.asarray(ps["species_weights"])[sidx].reshape(-1, 1)
            ps.update(positions=hx0, velocities=hv0, weights=w, species_integer_index=sidx,
                      charges=ps["charge_integer_lookup"][sidx].reshape(-1, 1) * w, masses=ps["mass_integer_lookup"][sidx].reshape(-1, 1) * w,
                      charge_to_mass_ratios=ps["charge_mass_integer_lookup"][sidx].reshape(-1, 1))
        else:
            ps.update(positions=None, velocities=None, weights=None, species_integer_index=None, charges=None, masses=None,
                      charge_to_mass_ratios=None)
        out_d = self._assemble_output(sec, state, ps, res, ext_E, ext_B)
        out_d["species_table"] = ps["species_table"]
        out_d["rank"], out_d["world_size"] = rank, world
        if "kinetic_energy_species" in res:
            out_d["kinetic_energy_species"] = res["kinetic_energy_species"]
        return out_d

    # ---- run ------------------------------------------------------------------------------------------------------
    def simulation(self, input_parameters=None):
        from ._engine import simulate_host
        sec = self._sections(self.clean_runtime_input_parameters(input_parameters))
        dom, solver, ext = sec["domain_parameters"], sec["solver_parameters"], sec["external_field_parameters"]
        state = self.build_domain_state(dom)
        rank, world = self._distributed()
        n_all = sum(sp["number_pseudoparticles"] for kind in sec["species_parameters"].values() for sp in kind.values())
        mode = solver.get("device_resident", "auto")
        can = (not bool(solver["particle_history"])) and solver.get("rng", "threefry") == "threefry" and solver["engine"] != "indexed"
        if mode is True and not can:
            raise JicError("device_resident=True needs particle_history=False, rng='threefry' and engine 'auto' or 'binned'")
        if world > 1 and not can:
            raise JicError("a run over several ranks needs particle_history=False, rng='threefry' and engine 'auto' or 'binned'")
        if can and (mode is True or (mode == "auto" and (n_all >= 1_000_000 or world > 1))):
            return self._simulation_device(sec, state, rank, world)
        ps = self.initialize_particle_state(sec["species_parameters"], dom, solver, state)
        if solver["print_info"]:  # the reference's start-up summary (_simulation.py:176-183)
            print(INFORMATION_TEXT.format(*simulation_information(dom, sec["species_parameters"], ext, state, ps)))
        G, T = int(dom["number_grid_points"]), int(dom["total_steps"])
        N = len(ps["positions"])
        dtype = np.float64 if str(solver["dtype"]) in ("float64", "f64") else np.float32
        eE = ext.get("external_electric_field"); eB = ext.get("external_magnetic_field")
        ext_E = np.asarray(eE["E"], np.float32) if isinstance(eE, dict) and "E" in eE else np.zeros((G, 3), np.float32)
        ext_B = np.asarray(eB["B"], np.float32) if isinstance(eB, dict) and "B" in eB else np.zeros((G, 3), np.float32)
        if ext_E.shape != (G, 3) or ext_B.shape != (G, 3):
            raise JicError(f"external fields must have shape ({G}, 3), got {ext_E.shape} and {ext_B.shape}")
        history = bool(solver["particle_history"])
        engine = solver["engine"]
        if engine == "auto":
            engine = "indexed" if history or N < 1_000_000 else "binned"
        if history and engine == "binned":
            raise JicError("particle_history=True needs engine='indexed' (the binned store does not keep particle order)")
        res = simulate_host(species=ps["species_table"], x0=ps["positions"], v0=ps["velocities"], n_steps=T, ext_E=ext_E, ext_B=ext_B,
                            dtype=dtype, fields=True, particles=history, initial=True, engine=engine,
                            length=state["box_size"][0], length_y=state["box_size"][1], length_z=state["box_size"][2], G=G, dt=state["dt"],
                            pbl=dom["particle_BC_left"], pbr=dom["particle_BC_right"], fbl=dom["field_BC_left"], fbr=dom["field_BC_right"],
                            filter_passes=solver["filter_passes"], filter_alpha=solver["filter_alpha"], filter_strides=solver["filter_strides"],
                            relativistic=bool(solver["relativistic"]), track_yz=history, field_solver=int(solver["field_solver"]),
                            time_evolution_algorithm=int(solver["time_evolution_algorithm"]),
                            cn_substeps=int(solver["number_of_particle_substeps_implicit_CN"]),
                            cn_max_iterations=int(solver["max_number_of_Picard_iterations_implicit_CN"]),
                            cn_tolerance=float(solver["tolerance_Picard_iterations_implicit_CN"]))
        return self._assemble_output(sec, state, ps, res, ext_E, ext_B)

    @staticmethod
    def _assemble_output(sec, state, ps, res, ext_E, ext_B):
        """jaxincell/_simulation.py:260-344: the output dictionary (`res` holds what the library produced)."""
        dom, solver, ext = sec["domain_parameters"], sec["solver_parameters"], sec["external_field_parameters"]
        G, T = int(dom["number_grid_points"]), int(dom["total_steps"])
        e0 = next(iter(sec["species_parameters"]["electrons"].values()))
        we = ps["species_weights"][0] if "species_weights" in ps else ps["weights"][0, 0]
        plasma_frequency = (np.sqrt(e0["number_pseudoparticles"] * we * ps["charge_electrons"] ** 2) / np.sqrt(mass_electron)
                            / np.sqrt(epsilon_0) / np.sqrt(state["box_size"][0]))
        out = {  # jaxincell/_simulation.py:269-312
            "positions": res.get("positions"), "velocities": res.get("velocities"), "masses": ps["masses"], "charges": ps["charges"],
            "charge_to_mass_ratios": ps["charge_to_mass_ratios"], "initial_positions": ps["positions"],
            "initial_velocities": (res["initial_velocities"] if "initial_velocities" in res else
                                   None if ps["velocities"] is None else
                                   post_bc_initial_velocities(ps["positions"], ps["velocities"], state["dt"], state["box_size"][0],
                                                              dom["particle_BC_left"], dom["particle_BC_right"])),
            "weights": ps["weights"], "species_integer_index": ps["species_integer_index"],
            "charge_integer_lookup": ps["charge_integer_lookup"], "mass_integer_lookup": ps["mass_integer_lookup"],
            "charge_mass_integer_lookup": ps["charge_mass_integer_lookup"],
            "electric_field": res["electric_field"], "magnetic_field": res["magnetic_field"], "current_density": res["current_density"],
            "charge_density": res["charge_density"], "number_grid_points": G, "number_pseudoelectrons": e0["number_pseudoparticles"],
            "total_steps": T, "time_array": np.linspace(0, T * state["dt"], T), "grid": state["grid"], "dt": state["dt"],
            "plasma_frequency": plasma_frequency, "max_initial_vth_electrons": ps["vth_electrons"],
            "vth_electrons_over_c": ps["vth_electrons_over_c"], "charge_electrons": ps["charge_electrons"], "dx": state["dx"],
            "length": state["box_size"][0], "box_size": state["box_size"],
            "fields": res["fields"], "external_electric_field": ext_E, "external_magnetic_field": ext_B,
        }
        return {**dom, **ext, **sec["source_parameters"], **solver, **out, "domain_parameters": dom,
                "species_parameters": sec["species_parameters"], "external_field_parameters": ext,
                "source_parameters": sec["source_parameters"], "solver_parameters": solver, "parameter_sections": sec}

    def output_keys(self, input_parameters=None):
        """Key set of the dictionary `run()` returns (host logic only: no particles are pushed)."""
        sec = self._sections(input_parameters)
        dom = sec["domain_parameters"]
        state = self.build_domain_state(dom)
        n = sum(sp["number_pseudoparticles"] for kind in sec["species_parameters"].values() for sp in kind.values())
        e0 = next(iter(sec["species_parameters"]["electrons"].values()))
        ps = dict(masses=None, charges=None, charge_to_mass_ratios=None, positions=None, velocities=None, weights=np.ones((n, 1)),
                  species_integer_index=None, charge_integer_lookup=None, mass_integer_lookup=None, charge_mass_integer_lookup=None,
                  vth_electrons=None, vth_electrons_over_c=None, charge_electrons=e0["charge_over_elementary_charge"] * elementary_charge)
        res = dict.fromkeys(("positions", "velocities", "electric_field", "magnetic_field", "current_density", "charge_density", "fields"))
        return sorted(self._assemble_output(sec, state, ps, res, None, None))

    def run(self, input_parameters=None):
        return self.simulation(input_parameters)


INFORMATION_TEXT = (  # jaxincell/_state_initialization.py:320-336
    "Length of the simulation box: {} Debye lengths or {} Skin Depths\n"
    "Density of electrons: {} m^-3\n"
    "Electron temperature: {} eV\n"
    "Ion temperature / Electron temperature: {}\n"
    "Debye length: {} m\n"
    "Skin depth: {} m\n"
    "Wavenumber * Debye length: {}\n"
    "Pseudoparticles per cell: {}\n"
    "Pseudoparticle weight: {}\n"
    "Steps at each plasma frequency: {}\n"
    "Total time: {} / plasma frequency\n"
    "Number of particles on a Debye cube: {}\n"
    "Relativistic gamma factor: Maximum {}, Average {}\n"
    "Charge x External electric field x Debye Length / Temperature: {}\n")


def simulation_information(dom, species_parameters, ext, state, ps):
    """The sixteen numbers of the reference's start-up summary (print_simulation_information, _state_initialization.py:288-357), in the
    order of INFORMATION_TEXT."""
    e0 = next(iter(species_parameters["electrons"].values()))
    i0 = next(iter(species_parameters["ions"].values()))
    length, dx, dt = state["box_size"][0], state["dx"], state["dt"]
    T, G, n_e = dom["total_steps"], dom["number_grid_points"], e0["number_pseudoparticles"]
    weight, q_e, vth = ps["weights"][0, 0], ps["charge_electrons"], ps["vth_electrons"]
    debye_per_dx = 1 / e0["grid_points_per_Debye_length"]
    temperature = mass_electron * vth ** 2 / 2 / (-q_e)
    plasma_frequency = np.sqrt(n_e * weight * q_e ** 2) / np.sqrt(mass_electron) / np.sqrt(epsilon_0) / np.sqrt(length)
    with np.errstate(divide="ignore", invalid="ignore"):
        gamma = 1 / np.sqrt(1 - np.sum(np.asarray(ps["velocities"]) ** 2, axis=1) / speed_of_light ** 2)
        field_term = -q_e * ext["external_electric_field_amplitude"] * debye_per_dx * dx / (mass_electron * vth ** 2 / 2)
    return [length / (debye_per_dx * dx), length / (speed_of_light / plasma_frequency), n_e * weight / length, temperature,
            i0["ion_temperature_over_electron_temperature_x"], debye_per_dx * dx, speed_of_light / plasma_frequency,
            e0["perturbation_wavenumber_x"] * debye_per_dx * dx, n_e / G, weight, 1 / (plasma_frequency * dt), dt * plasma_frequency * T,
            n_e * weight / length * (debye_per_dx * dx) ** 3, np.max(gamma), np.mean(gamma), field_term]


def simulation(parameters=None, input_parameters=None):
    """Free-function spelling of the entry point (BASELINE.json calls it `simulation()`; this version of the reference spells it
    `Simulation(parameters).run(input_parameters)`, jaxincell/_simulation.py:94-121): parameters is a dict or the path of a TOML file."""
    return Simulation(parameters).run(input_parameters)


def _field_energies(output):
    """The field part of jaxincell/_diagnostics.py:98-129 (needs no particles)."""
    T, dt, dx = int(output["total_steps"]), float(output["dt"]), output["dx"]
    sig = output["electric_field"][:, len(output["grid"]) // 2, 0]
    sig = (sig - np.mean(sig)) / np.max(sig)
    half = T // 2
    mag = np.abs(np.fft.fft(sig)[:half])
    freqs = np.fft.fftfreq(T, d=dt)[:half] * 2 * np.pi
    dominant = np.abs(freqs[np.argmax(mag)])
    E2 = np.sum(output["electric_field"] ** 2, axis=-1); B2 = np.sum(output["magnetic_field"] ** 2, axis=-1)
    eE2 = np.sum(np.asarray(output["external_electric_field"], np.float64) ** 2, axis=-1)
    eB2 = np.sum(np.asarray(output["external_magnetic_field"], np.float64) ** 2, axis=-1)
    return {
        "electric_field_energy_density": epsilon_0 / 2 * E2, "electric_field_energy": epsilon_0 / 2 * np.sum(E2, axis=-1) * dx,
        "magnetic_field_energy_density": 1 / (2 * mu_0) * B2, "magnetic_field_energy": 1 / (2 * mu_0) * np.sum(B2, axis=-1) * dx,
        "dominant_frequency": dominant,
        "external_electric_field_energy_density": epsilon_0 / 2 * eE2, "external_electric_field_energy": epsilon_0 / 2 * np.sum(eE2) * dx,
        "external_magnetic_field_energy_density": 1 / (2 * mu_0) * eB2, "external_magnetic_field_energy": 1 / (2 * mu_0) * np.sum(eB2) * dx,
    }


def _diagnostics_from_energy_history(output):
    """diagnostics() for a device-resident run: the same energies (jaxincell/_diagnostics.py:98-146), the kinetic ones from the
    per-species history the library reduced on the device instead of from a (T,N,3) velocity history that was never stored.
    Electrons = the species with negative charge, ions = the others (the reference's split by the sign of the charge, :31-32)."""
    ke = np.asarray(output["kinetic_energy_species"], dtype=np.float64)
    q = np.array([s["q"] for s in output["species_table"]])
    ke_e, ke_i = ke[:, q < 0].sum(axis=1), ke[:, q >= 0].sum(axis=1)
    output.update(_field_energies(output))
    output.update({"kinetic_energy": ke_e + ke_i, "kinetic_energy_electrons": ke_e, "kinetic_energy_ions": ke_i,
                   "species": [dict(name=s.get("name", f"species_{i}"), charge=float(s["q"]), mass=float(s["m"]), kinetic_energy=ke[:, i])
                               for i, s in enumerate(output["species_table"])]})
    output["total_energy"] = (output["electric_field_energy"] + output["external_electric_field_energy"] + output["magnetic_field_energy"]
                              + output["external_magnetic_field_energy"] + output["kinetic_energy"])
    return output


def post_bc_initial_velocities(x0, v0, dt, length, pbl, pbr):
    """The reference's output key `initial_velocities` (jaxincell/_simulation.py:217,286): the velocities after set_BC_particles was
    applied to the start-up half step x0 + dt/2 v0 -- v_x flipped for a particle a reflective wall sent back, zero for an absorbed one
    (_boundary_conditions.py:32-56).  Used when the library cannot return them (the binned store keeps no particle order)."""
    x0, v0 = np.asarray(x0), np.array(v0, copy=True)
    xp = x0[:, 0] + 0.5 * dt * v0[:, 0]
    for out_side, bc in ((xp < -length / 2, pbl), (xp > length / 2, pbr)):
        if bc == 1:
            v0[out_side, 0] = -v0[out_side, 0]
        elif bc == 2:
            v0[out_side, :] = 0.0
    return v0


def diagnostics(output):
    """jaxincell/_diagnostics.py:8-147 in NumPy: species split, energies, dominant frequency.  Mutates and returns `output`."""
    if output.get("positions") is None or output.get("velocities") is None:
        if output.get("kinetic_energy_species") is None:
            raise JicError("diagnostics() needs the particle histories (solver_parameters['particle_history']=True) or, for runs too "
                           "large for them, the device-reduced kinetic energies (solver_parameters['kinetic_energy_history']=True)")
        return _diagnostics_from_energy_history(output)
    q = np.asarray(output["charges"]).reshape(-1); m = np.asarray(output["masses"]).reshape(-1)
    esel, isel = q < 0, q >= 0
    output.update(position_electrons=output["positions"][:, esel, :], velocity_electrons=output["velocities"][:, esel, :],
                  mass_electrons=output["masses"][esel], charge_electrons=output["charges"][esel],
                  position_ions=output["positions"][:, isel, :], velocity_ions=output["velocities"][:, isel, :],
                  mass_ions=output["masses"][isel], charge_ions=output["charges"][isel])
    pairs, labels = np.unique(np.stack([q, m], axis=1), axis=0, return_inverse=True)
    labels = labels.reshape(-1)
    species = []
    for si, (qv, mv) in enumerate(pairs):
        mask = labels == si
        if qv < 0 and not any(s["name"] == "electrons" for s in species):
            name = "electrons"
        elif qv > 0 and not any(s["name"] == "ions" for s in species):
            name = "ions"
        else:
            name = f"species_{si}"
        species.append(dict(name=name, charge=float(qv), mass=float(mv), positions=output["positions"][:, mask, :],
                            velocities=output["velocities"][:, mask, :]))
    output["species"] = species
    for k in ("positions", "velocities", "masses", "charges"):
        del output[k]
    T, dt, dx = int(output["total_steps"]), float(output["dt"]), output["dx"]
    sig = output["electric_field"][:, len(output["grid"]) // 2, 0]
    sig = (sig - np.mean(sig)) / np.max(sig)
    half = T // 2
    mag = np.abs(np.fft.fft(sig)[:half])
    freqs = np.fft.fftfreq(T, d=dt)[:half] * 2 * np.pi
    dominant = np.abs(freqs[np.argmax(mag)])
    E2 = np.sum(output["electric_field"] ** 2, axis=-1); B2 = np.sum(output["magnetic_field"] ** 2, axis=-1)
    eE2 = np.sum(np.asarray(output["external_electric_field"], np.float64) ** 2, axis=-1)
    eB2 = np.sum(np.asarray(output["external_magnetic_field"], np.float64) ** 2, axis=-1)
    ke_e = 0.5 * np.sum(output["mass_electrons"].reshape(-1) * np.sum(output["velocity_electrons"] ** 2, axis=-1), axis=-1)
    ke_i = 0.5 * np.sum(output["mass_ions"].reshape(-1) * np.sum(output["velocity_ions"] ** 2, axis=-1), axis=-1)
    output.update({
        "electric_field_energy_density": epsilon_0 / 2 * E2, "electric_field_energy": epsilon_0 / 2 * np.sum(E2, axis=-1) * dx,
        "magnetic_field_energy_density": 1 / (2 * mu_0) * B2, "magnetic_field_energy": 1 / (2 * mu_0) * np.sum(B2, axis=-1) * dx,
        "dominant_frequency": dominant, "kinetic_energy": ke_e + ke_i, "kinetic_energy_electrons": ke_e, "kinetic_energy_ions": ke_i,
        "external_electric_field_energy_density": epsilon_0 / 2 * eE2, "external_electric_field_energy": epsilon_0 / 2 * np.sum(eE2) * dx,
        "external_magnetic_field_energy_density": 1 / (2 * mu_0) * eB2, "external_magnetic_field_energy": 1 / (2 * mu_0) * np.sum(eB2) * dx,
    })
    output["total_energy"] = (output["electric_field_energy"] + output["external_electric_field_energy"] + output["magnetic_field_energy"]
                              + output["external_magnetic_field_energy"] + output["kinetic_energy"])
    return output
