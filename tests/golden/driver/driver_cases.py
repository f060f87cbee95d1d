"""Parameter dictionaries shared by tests/golden/driver/make_reference_driver_golden.py (fed to the reference's `Simulation`) and
tests/test_simulation_driver.py (fed to jaxincell_b200.Simulation).  Written for this repository (shapes follow the reference's
examples: two-stream with an ion species that references the electrons, a four-species beam/neutraliser set-up with
`seed_position_override`, a Landau-type perturbation, walls); TEST INFRASTRUCTURE."""

CASES = {
    "defaults": {},
    "two_stream_references": {
        "domain_parameters": {"length": 0.01, "timestep_over_spatialstep_times_c": 4.5, "number_grid_points": 70, "number_grid_points_y": 3,
                              "number_grid_points_z": 3, "total_steps": 40},
        "solver_parameters": {"print_info": False, "tolerance_Picard_iterations_implicit_CN": 1e-10},
        "species_parameters": {
            "electrons": {"electrons0": {"number_pseudoparticles": 350, "grid_points_per_Debye_length": 0.50265482457, "perturbation_amplitude_x": 5e-7,
                                         "perturbation_wavenumber_x": 1, "vth_over_c_x": 0.05, "drift_speed_x": 6e7, "velocity_plus_minus_x": True}},
            "ions": {"ions0": {"number_pseudoparticles": 350, "grid_points_per_Debye_length": 0.50265482457, "vth_over_c_x": "_electrons0",
                               "vth_over_c_y": "_electrons0", "vth_over_c_z": "_electrons0", "ion_temperature_over_electron_temperature_x": 0.01}},
        },
    },
    "beam_four_species": {
        "domain_parameters": {"length": 1.0, "timestep_over_spatialstep_times_c": 3.0, "number_grid_points": 24, "total_steps": 6},
        "solver_parameters": {"print_info": False, "seed": 250724, "filter_passes": 0},
        "species_parameters": {
            "electrons": {
                "electrons0": {"number_pseudoparticles": 120, "grid_points_per_Debye_length": 2.565, "random_positions_x": True,
                               "random_positions_y": False, "random_positions_z": False, "vth_over_c_x": 0.07071067812, "drift_speed_x": -2.25e6},
                "beam": {"number_pseudoparticles": 90, "grid_points_per_Debye_length": 0.44427103214, "random_positions_x": True,
                         "vth_over_c_x": 0.07071067812, "drift_speed_x": 7.5e7, "seed_position_override": True, "seed_position": 10},
            },
            "ions": {
                "ions0": {"number_pseudoparticles": 120, "grid_points_per_Debye_length": 2.565, "random_positions_x": True,
                          "mass_over_proton_mass": 2.0, "vth_over_c_x": "_electrons0", "vth_over_c_y": "_electrons0", "vth_over_c_z": "_electrons0"},
                "beam_neutralizer": {"number_pseudoparticles": 90, "grid_points_per_Debye_length": "_electrons1", "random_positions_x": True,
                                     "vth_over_c_x": 0.0016, "seed_position_override": True, "seed_position": 10},
                "heavy": {"number_pseudoparticles": 40, "weight": 3.5e9, "charge_over_elementary_charge": 2, "mass_over_proton_mass": 4.0,
                          "vth_over_c_y": "_ions0", "drift_speed_y": 1e4, "velocity_plus_minus_y": True},
            },
        },
    },
    "landau_walls_relativistic": {
        "domain_parameters": {"length": 1.0, "length_y": 0.5, "length_z": 0.25, "timestep_over_spatialstep_times_c": 1.0, "number_grid_points": 32,
                              "total_steps": 8, "particle_BC_left": 1, "particle_BC_right": 2, "field_BC_left": 1, "field_BC_right": 2},
        "solver_parameters": {"print_info": False, "relativistic": True, "filter_passes": 3, "filter_alpha": 0.4, "filter_strides": [1, 3], "seed": 7},
        "species_parameters": {
            "electrons": {"number_pseudoparticles": 200, "perturbation_amplitude_x": 0.025, "perturbation_wavenumber_x": 1.02, "vth_over_c_x": 0.35,
                          "vth_over_c_z": 0.1, "drift_speed_x": 0, "velocity_plus_minus_x": False, "random_positions_x": False},
            # (random x: cold ions on a linspace would sit exactly ON the walls, where one ulp decides between "inside" and "absorbed")
            "ions": {"number_pseudoparticles": 160, "mass_over_proton_mass": 1e9, "vth_over_c_x": 0.0, "vth_over_c_y": 0.0, "vth_over_c_z": 0.0,
                     "perturbation_amplitude_x": 0.0, "random_positions_x": True},
        },
    },
}

# case run end to end by the reference (output-dict contract, plasma frequency, diagnostics); explicit particles are not needed:
# the draws come from the Threefry restatement on both sides (oracle/sampling.py under the stand-in, jic_sample_particles on the GPU)
RUN_CASE = "two_stream_references"
# every case the reference also RUNS (histories stored); the first one additionally feeds the diagnostics comparison
RUN_CASES = ("two_stream_references", "beam_four_species", "landau_walls_relativistic")

# runtime overrides passed to `run(input_parameters)` / `clean_runtime_input_parameters` on top of a case (the reference only admits its
# differentiable parameters there, _routing.py:160-226): the electron thermal speed changes, so the ions' "_electrons0" reference must follow
RUNTIME_OVERRIDES = {
    "two_stream_references": {"length": 0.02, "filter_alpha": 0.3, "timestep_over_spatialstep_times_c": 2.0,
                              "electrons": {"electrons0": {"vth_over_c_x": 0.08, "drift_speed_x": 5e7}},
                              "ions": {"mass_over_proton_mass": 2.0}},
    "beam_four_species": {"ions": {"_ions2": {"weight": 1.0e9, "drift_speed_y": -2e4}, "beam_neutralizer": {"vth_over_c_x": 0.002}},
                          "electrons": {"grid_points_per_Debye_length": 1.5}},
}
RUNTIME_INVALID = [  # (case, input, exception, text the message must contain)
    ("two_stream_references", {"total_steps": 2}, "ValueError", "total_steps"),
    ("two_stream_references", {"ions": {"ions0": {"number_pseudoparticles": 3}}}, "ValueError", "ions.ions0.number_pseudoparticles"),
    ("two_stream_references", {"ion_drift_speed_x": 3.0}, "ValueError", "ion_drift_speed_x"),
    ("two_stream_references", {"electrons": {"nobody": {"vth_over_c_x": 0.1}}}, "ValueError", "nobody"),
    ("two_stream_references", [("length", 1.0)], "TypeError", "dictionary"),
]
