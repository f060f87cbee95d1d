"""The device-resident path of the drop-in driver (`Simulation.run()` with particle_history=False): particles sampled on the GPU and
handed to the library there, index-sharded over the ranks of a torch.distributed group, per-species kinetic energies reduced on the
device.  Reference behaviour being reproduced: jaxincell/_simulation.py:94-121,169-190,256-312 and _diagnostics.py:98-146."""
import copy
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FIELDS = ("electric_field", "magnetic_field", "current_density", "charge_density")


def _parameters(n=40_000, G=128, T=30, **solver):
    return {
        "domain_parameters": dict(total_steps=T, number_grid_points=G, length=0.02, timestep_over_spatialstep_times_c=0.8),
        "species_parameters": {
            "electrons": dict(number_pseudoparticles=n, vth_over_c_x=0.05, vth_over_c_y=0.01, vth_over_c_z=0.01, random_positions_x=True,
                              velocity_plus_minus_x=True, drift_speed_x=4e7, perturbation_amplitude_x=1e-4, perturbation_wavenumber_x=2),
            "ions": dict(number_pseudoparticles=n, random_positions_x=True, ion_temperature_over_electron_temperature_x=0.01),
        },
        "solver_parameters": dict(print_info=False, particle_history=False, **solver),
    }


def _rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))


def test_device_resident_run_equals_the_host_path():
    """Same dictionary through both paths: the host path samples on the device too but stages everything through NumPy and steps with
    the INDEXED engine; the device-resident path never leaves the GPU and steps with the binned engine."""
    from jaxincell_b200 import Simulation
    host = Simulation(_parameters(device_resident=False)).run()
    dev = Simulation(_parameters(device_resident=True)).run()
    assert set(host) <= set(dev)
    for k in FIELDS:
        assert _rel(dev[k], host[k]) < 1e-8, k
    np.testing.assert_array_equal(dev["initial_positions"], host["initial_positions"])
    np.testing.assert_array_equal(dev["initial_velocities"], host["initial_velocities"])
    for k in ("weights", "charges", "masses", "charge_to_mass_ratios", "species_integer_index"):
        np.testing.assert_array_equal(dev[k], host[k])
    for k in ("plasma_frequency", "dt", "dx"):
        assert dev[k] == host[k]
    assert _rel(dev["fields"][0], host["fields"][0]) < 1e-12  # (rho_0 summed in a different order by the two engines)


def test_kinetic_energy_history_matches_the_particle_histories():
    """diagnostics() of a run that kept (T,N,3) velocities (INDEXED engine) vs diagnostics() of the device-resident run of the same
    dictionary, whose kinetic energies were reduced on the device step by step."""
    from jaxincell_b200 import Simulation, diagnostics
    par = _parameters(n=20_000, T=24)
    par["solver_parameters"].update(particle_history=True)
    full = diagnostics(Simulation(par).run())
    dev = diagnostics(Simulation(_parameters(n=20_000, T=24, device_resident=True, kinetic_energy_history=True)).run())
    for k in ("kinetic_energy", "kinetic_energy_electrons", "kinetic_energy_ions", "electric_field_energy", "magnetic_field_energy", "total_energy"):
        assert _rel(dev[k], full[k]) < 1e-8, k
    assert dev["dominant_frequency"] == pytest.approx(full["dominant_frequency"], rel=1e-12)


def test_crank_nicolson_device_resident_takes_the_sorted_push_and_matches_the_particle_histories():
    """time_evolution_algorithm = 1 with 2.4e5 particles: above the library's threshold, so the device-resident run steps with the
    cell-sorted Crank-Nicolson push; its fields and device-reduced kinetic energies against the host path of the same dictionary with
    JIC_CN_SORTED_MIN raised (unsorted push, (T,N,3) histories kept, energies from the velocity history)."""
    import os
    from jaxincell_b200 import Simulation, diagnostics
    cn = dict(time_evolution_algorithm=1, max_number_of_Picard_iterations_implicit_CN=8, tolerance_Picard_iterations_implicit_CN=1e-9,
              number_of_particle_substeps_implicit_CN=2)
    par = _parameters(n=120_000, T=10, **cn)
    par["solver_parameters"].update(particle_history=True)
    os.environ["JIC_CN_SORTED_MIN"] = str(1 << 40)
    try:
        full = diagnostics(Simulation(par).run())
    finally:
        del os.environ["JIC_CN_SORTED_MIN"]
    dev = diagnostics(Simulation(_parameters(n=120_000, T=10, device_resident=True, kinetic_energy_history=True, **cn)).run())
    for k in FIELDS:
        assert _rel(dev[k], full[k]) < 1e-8, k
    for k in ("kinetic_energy", "kinetic_energy_electrons", "kinetic_energy_ions", "total_energy"):
        assert _rel(dev[k], full[k]) < 1e-8, k


def test_kinetic_energy_history_through_the_c_abi_all_engines():
    """jic_outputs.kinetic_energy on both engines against 0.5 m v^2 of the velocity history (INDEXED keeps it)."""
    import torch
    from jaxincell_b200 import HotPath
    from plasma import cfl_dt, two_species
    G, length, T = 48, 0.01, 10
    p = two_species(9000, 7000, length=length, G=G, seed=3, vth_e=0.1, vth_yz=0.05, drift=3e7, plus_minus=True, gpdl=0.5)
    dt = cfl_dt(length, G, 0.9)
    got = {}
    for engine in ("indexed", "binned"):
        hp = HotPath(species=p["species"], length=length, G=G, dt=dt, engine=engine, track_yz=engine == "indexed")
        hp.set_external_fields(None, None)
        hp.initialize(p["x0"], p["v0"])
        out = hp.run(T, particles=engine == "indexed", kinetic=True)
        hp.check_status()
        got[engine] = out["kinetic_energy"].cpu().numpy()
        if engine == "indexed":
            v = out["velocities"].double().cpu().numpy()
            m = p["m"]
            ne = p["species"][0]["count"]
            want = np.stack([0.5 * (m[:ne] * (v[:, :ne] ** 2).sum(-1)).sum(-1), 0.5 * (m[ne:] * (v[:, ne:] ** 2).sum(-1)).sum(-1)], axis=1)
            assert _rel(got[engine], want) < 1e-12
        hp.close()
    assert _rel(got["binned"], got["indexed"]) < 1e-9


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from jaxincell_b200 import Simulation, diagnostics
        out = diagnostics(Simulation(_parameters(n=30_001, G=512, T=16, kinetic_energy_history=True)).run())
        q.put((rank, {k: np.asarray(out[k]) for k in FIELDS + ("kinetic_energy", "initial_positions")}, out["world_size"]))
    except Exception as e:  # noqa: BLE001
        q.put((rank, f"{type(e).__name__}: {e}", None))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 4])
def test_simulation_under_torch_distributed_shards_by_index(world):
    """`Simulation(parameters).run()` inside an NCCL process group: every rank samples and pushes its index slice only; fields and
    kinetic energies equal the single-rank run of the same dictionary, the per-particle arrays are the rank's slices."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from jaxincell_b200 import Simulation, diagnostics
    from jaxincell_b200._parallel import shard_counts
    one = diagnostics(Simulation(_parameters(n=30_001, G=512, T=16, device_resident=True, kinetic_energy_history=True)).run())
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = sorted((q.get(timeout=240) for _ in procs), key=lambda t: t[0])
    for pr in procs:
        pr.join(timeout=60)
    pos = []
    for rank, out, ws in res:
        assert isinstance(out, dict), out
        assert ws == world
        for k in FIELDS + ("kinetic_energy",):
            assert _rel(out[k], one[k]) < 1e-8, (rank, k)
        np.testing.assert_array_equal(out["electric_field"], res[0][1]["electric_field"])  # bit-identical across ranks
        pos.append(out["initial_positions"])
    # the slices tile the single-rank arrays: species block by species block
    n = 30_001
    counts = shard_counts(n, world)
    for s in range(2):
        parts = []
        for r in range(world):
            lo = 0 if s == 0 else counts[r]
            parts.append(pos[r][lo:lo + counts[r]])
        np.testing.assert_array_equal(np.concatenate(parts), one["initial_positions"][s * n:(s + 1) * n])
