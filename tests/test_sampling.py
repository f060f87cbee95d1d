"""Initial particle sampling: the jax.random restatement in oracle/sampling.py against published
known-answer values, and the device sampler (jic_sample_particles) against the oracle.

Reference: jaxincell/_state_initialization.py:51-96,259-260; its own tests only compare against jax.random itself
(tests/test_state_initialization.py:81-97,422-570), so the vectors below come from the Random123 distribution (Threefry-2x32-20
known-answer file, also used by jax's own test suite) and from the scalar draws printed in the JAX documentation."""
import numpy as np
import pytest

from oracle import sampling as S

SPECIES = [
    dict(count=5001, seed_position=1701, seed_velocity=1704, random_positions=[False, True, True], velocity_plus_minus=[True, False, False],
         perturbation_amplitude=[1e-7, 0.0, 0.0], perturbation_wavenumber=[8, 0, 0], vth_over_c=[0.05, 0.0, 0.01], drift_speed=[1e8, 0.0, 0.0]),
    dict(count=3000, seed_position=1701, seed_velocity=1707, random_positions=[True, True, True], velocity_plus_minus=[False, False, True],
         perturbation_amplitude=[0.0, 2e-4, 0.0], perturbation_wavenumber=[0, 3, 0], vth_over_c=[0.9, 1e-3, 1e-3], drift_speed=[0.0, 5e6, 0.0]),
    dict(count=1, seed_position=1713, seed_velocity=1713, random_positions=[False, False, True], velocity_plus_minus=[False, False, False],
         perturbation_amplitude=[0.0, 0.0, 0.0], perturbation_wavenumber=[0, 0, 0], vth_over_c=[0.1, 0.1, 0.1], drift_speed=[0.0, 0.0, 0.0]),
]
BOX = (0.01, 0.02, 0.005)


@pytest.mark.parametrize("key,ctr,want", [
    ((0x00000000, 0x00000000), (0x00000000, 0x00000000), (0x6b200159, 0x99ba4efe)),
    ((0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff), (0x1cb996fc, 0xbb002be7)),
    ((0x13198a2e, 0x03707344), (0x243f6a88, 0x85a308d3), (0xc4923a9c, 0x483df7a0)),
])
def test_threefry2x32_random123_known_answers(key, ctr, want):
    y0, y1 = S.threefry2x32(key[0], key[1], ctr[0], ctr[1])
    assert (int(y0), int(y1)) == want


def test_scalar_draws_quoted_in_the_jax_documentation():
    assert abs(S.uniform32_scalar(0, partitionable=False) - 0.41845703) < 1e-8
    assert abs(S.uniform32_scalar(0, partitionable=True) - 0.947667) < 1e-6
    assert abs(S.normal32_scalar(0, partitionable=False) - (-0.20584226)) < 2e-7
    assert abs(S.normal32_scalar(42, partitionable=False) - (-0.18471177)) < 2e-7
    assert abs(S.normal32_scalar(42, partitionable=True) - (-0.028304616)) < 2e-7


def test_vector_draws_are_distributed_as_documented():
    u = S.uniform64(1702, 200_000, -0.005, 0.005)
    assert u.min() >= -0.005 and u.max() < 0.005
    assert abs(u.mean()) < 5e-5 and abs(u.std() - 0.01 / np.sqrt(12)) < 2e-5
    z = S.normal64(1705, 200_000)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01 and abs((z ** 3).mean()) < 0.03
    for part in (True, False):
        a, b = S.random_bits64(7, 64, part), S.random_bits64(7, 65, part)
        assert (a[:10] == b[:10]).all() == part  # the original layout depends on the draw length, the partitionable one does not
    assert len(np.unique(S.random_bits64(3, 50_000))) == 50_000


def test_species_phase_space_formulas():
    x, v = S.sample(SPECIES, BOX)
    n0 = SPECIES[0]["count"]
    lin = S.jnp_linspace(-BOX[0] / 2, BOX[0] / 2, n0)
    assert lin[0] == -BOX[0] / 2 and lin[-1] == BOX[0] / 2
    np.testing.assert_allclose(x[:n0, 0], lin + 1e-7 * np.sin(8 * 2 * np.pi / BOX[0] * lin), rtol=0, atol=1e-18)
    assert (np.sign(v[:n0:2, 0]) > 0).mean() > 0.99 and (np.sign(v[1:n0:2, 0]) < 0).mean() > 0.99   # (-1)**arange
    assert (v[:n0, 1] == 0).all()
    assert np.abs(v).max() == 0.99 * S.speed_of_light                                                    # vth 0.9c species gets clipped
    assert S.species_seed_pair(1701, "electrons", 0, 0) == (1701, 1704) and S.species_seed_pair(1701, "ions", 1, 1) == (1719, 1719)


@pytest.mark.gpu
@pytest.mark.parametrize("partitionable", [True, False])
def test_device_sampler_matches_the_oracle(partitionable):
    import torch
    from jaxincell_b200 import sample_particles
    x_ref, v_ref = S.sample(SPECIES, BOX, partitionable)
    x, v = sample_particles(SPECIES, BOX, threefry_partitionable=partitionable)
    x, v = x.cpu().numpy(), v.cpu().numpy()
    # positions: integer Threefry + one multiply-add -> identical up to fma contraction; velocities: erfinv implementations differ by ulps
    np.testing.assert_allclose(x, x_ref, rtol=0, atol=4e-16 * max(BOX))
    np.testing.assert_allclose(v, v_ref, rtol=1e-12, atol=1e-12 * 0.05 * S.speed_of_light)
    x32, v32 = sample_particles(SPECIES, BOX, dtype=torch.float32, threefry_partitionable=partitionable)
    np.testing.assert_allclose(v32.cpu().numpy(), v_ref, rtol=1e-6, atol=1e-6 * 0.05 * S.speed_of_light)


@pytest.mark.gpu
def test_driver_uses_the_device_sampler_and_the_seed_schedule():
    from jaxincell_b200 import Simulation
    par = {"domain_parameters": {"number_grid_points": 16, "total_steps": 3, "length": 0.01},
           "species_parameters": {"electrons": {"e": {"number_pseudoparticles": 300, "vth_over_c_x": 0.05, "random_positions_x": True,
                                                       "perturbation_amplitude_x": 0.0, "drift_speed_x": 0.0, "velocity_plus_minus_x": False}},
                                  "ions": {"i": {"number_pseudoparticles": 200, "vth_over_c_x": "e"}}},
           "solver_parameters": {"print_info": False, "seed": 99}}
    sim = Simulation(par)
    st = sim.build_domain_state(sim.domain_parameters)
    ps = sim.initialize_particle_state(sim.species_parameters, sim.domain_parameters, sim.solver_parameters, st)
    np.testing.assert_allclose(ps["positions"][:300, 0], S.uniform64(99 + 1, 300, -0.005, 0.005), rtol=0, atol=1e-17)
    np.testing.assert_allclose(ps["positions"][300:, 1], S.uniform64(99 + 2, 200, -0.005, 0.005), rtol=0, atol=1e-17)  # ions share seed_position
    vth = 0.05 * S.speed_of_light / np.sqrt(2)
    np.testing.assert_allclose(ps["velocities"][:300, 0], vth * S.normal64(99 + 3 + 4, 300), rtol=1e-12, atol=1e-9)
    out = sim.run()
    np.testing.assert_array_equal(out["initial_positions"], ps["positions"])
