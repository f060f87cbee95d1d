"""GPU legs of the golden vectors that were added after this round's GPU budget was spent: every `refsrc_*` file (the reference's own
source on the NumPy stand-in, tests/golden/make_reference_golden.py) and the five newest cases of tests/golden/make_golden.py
(field_solver 3, field_solver 1 with walls, G = 5 and G = 3 grids, a box with its own length_y / length_z).

Same check as tests/test_golden.py::test_cuda_reproduces_golden (both CUDA engines through the C ABI, 1e-5 relative in fp64).  The file
sorts last on purpose: the driver runs `pytest -x`, these cases have not been on hardware yet, and a surprise here must not hide the
rest of the suite."""
import copy
import os

import numpy as np
import pytest

from test_golden import GPU_FILES_LATE, cuda_reproduces_golden
from test_simulation_driver import DRIVER_CASES, REFSRC, RUN_CASE, RUN_CASES, Simulation, _run_arrays


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["indexed", "binned"])
@pytest.mark.parametrize("path", GPU_FILES_LATE, ids=[os.path.basename(f)[:-4] for f in GPU_FILES_LATE])
def test_cuda_reproduces_late_golden(path, engine):
    cuda_reproduces_golden(path, engine)


@pytest.mark.gpu
@pytest.mark.parametrize("name", RUN_CASES)
def test_run_of_the_reference_source_is_reproduced_end_to_end(name):
    """`Simulation(parameters).run()` here vs the reference's own `Simulation(parameters).run()` (on the stand-in) for the same parameter
    dictionary -- initial particles from the device Threefry sampler, every history: the two-stream set-up at CFL 4.5 (multi-cell
    jumps), five species with cross references / seed overrides at CFL 3 without filter, walls + relativistic push."""
    ref = REFSRC[name]
    a, run = _run_arrays(name)
    out = Simulation(copy.deepcopy(DRIVER_CASES[name])).run()
    if name == RUN_CASE:
        assert set(ref["output_keys"]) <= set(out)
        np.testing.assert_allclose([out["time_array"][0], out["time_array"][1], out["time_array"][-1]], ref["time_array"][:3], rtol=1e-14)
    np.testing.assert_allclose(out["plasma_frequency"], ref["plasma_frequency"], rtol=1e-14)
    box = max(ref["box_size"])
    np.testing.assert_allclose(out["initial_positions"], a[f"{name}__positions"], rtol=0, atol=1e-15 * box)
    for k in ("electric_field", "magnetic_field", "current_density", "charge_density", "positions", "velocities"):
        err = np.abs(np.asarray(out[k]) - run[k]).max() / max(np.abs(run[k]).max(), 1e-300)
        assert err < 1e-5, (name, k, err)


@pytest.mark.gpu
@pytest.mark.parametrize("name", RUN_CASES)
def test_initial_fields_attribute_matches_the_reference_source(name):
    """`Simulation(parameters).fields` (E_x from Gauss's law on the filtered initial charge, _state_initialization.py:365-378) with the
    reference's own initial particles injected through `initial_positions` / `initial_velocities`."""
    a, _ = _run_arrays(name)
    par = copy.deepcopy(DRIVER_CASES[name])
    o = 0
    for kind in ("electrons", "ions"):
        group = par["species_parameters"][kind]
        group = group if any(isinstance(v, dict) for v in group.values()) else {None: group}
        for sp in group.values():
            n = sp["number_pseudoparticles"]
            sp["initial_positions"] = a[f"{name}__positions"][o:o + n]
            sp["initial_velocities"] = a[f"{name}__velocities"][o:o + n]
            o += n
    sim = Simulation(par)
    E0, B0 = sim.fields
    ref = a[f"{name}__E0"]
    assert np.abs(E0 - ref).max() < 1e-9 * np.abs(ref).max()
    assert not B0.any() and sim.external_electric_field.shape == ref.shape


# ---- direct per-step parity at BASELINE.json sizes, against the compiled oracle (oracle/c/jic_oracle.c) ---------------------------------
def _full_size_case(n_e, n_i, G, length, cfl, seed, vth, drift, plus_minus, ion_mass=1.0):
    from oracle import literal as L
    rng = np.random.default_rng(seed)
    c = L.speed_of_light
    N = n_e + n_i
    x0 = rng.uniform(-length / 2, length / 2, (N, 3))
    v0 = np.empty((N, 3))
    for a in range(3):
        v0[:n_e, a] = vth[a] * c / np.sqrt(2) * rng.standard_normal(n_e)
    v0[:n_e, 0] += drift
    if plus_minus:
        v0[1:n_e:2, 0] *= -1.0
    mi = ion_mass * L.mass_proton
    vthi = vth[0] * np.sqrt(L.mass_electron / mi)
    v0[n_e:] = vthi * c / np.sqrt(2) * rng.standard_normal((n_i, 3))
    np.clip(v0, -0.99 * c, 0.99 * c, out=v0)
    w = lambda n: L.epsilon_0 * L.mass_electron * c ** 2 / L.elementary_charge ** 2 * G ** 2 / length / (2 * n) * max(vth) ** 2 * 2.0 ** 2  # noqa: E731
    species = [dict(count=n_e, q=-L.elementary_charge * w(n_e), m=L.mass_electron * w(n_e), qm=-L.elementary_charge / L.mass_electron),
               dict(count=n_i, q=L.elementary_charge * w(n_i), m=mi * w(n_i), qm=L.elementary_charge / mi)]
    per = lambda key: np.concatenate([np.full(s["count"], s[key]) for s in species])  # noqa: E731
    return dict(x0=x0, v0=v0, q=per("q"), m=per("m"), qm=per("qm"), species=species, dt=cfl * (length / G) / c)


@pytest.mark.gpu
@pytest.mark.parametrize("name,n_e,n_i,G,cfl,vth,drift,pm,T", [
    ("two_stream_1e7_electrons", 10_000_000, 2_000_000, 4096, 1.0, (0.05, 0.0, 0.0), 6e7, True, 12),       # BASELINE.json config 2
    ("weibel_1d3v", 6_000_000, 6_000_000, 4096, 1.0, (0.01, 0.10, 0.10), 0.0, False, 10),                   # config 3 (reduced to 1.2e7)
    ("scaling_shape", 8_000_000, 8_000_000, 4096, 1.0, (0.05, 0.01, 0.01), 0.2 * 2.99792458e8, True, 10),   # config 5 shape at 1.6e7
])
def test_binned_engine_against_the_compiled_oracle_at_scale(name, n_e, n_i, G, cfl, vth, drift, pm, T):
    """Per-step E, B, J, rho of the BINNED engine (the bench's path) vs oracle/c/jic_oracle.c on identical particles, >= 1.2e7 of them:
    the same 1e-5 relative bound the north star states for small cases, now at sizes the NumPy oracle cannot reach in test time."""
    import torch
    from jaxincell_b200 import HotPath
    from oracle import c_port as CP
    length = G * 0.01 / 70
    p = _full_size_case(n_e, n_i, G, length, cfl, 1701, vth, drift, pm)
    ref = CP.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=p["dt"], total_steps=T, keep_particles=False)
    hp = HotPath(species=p["species"], length=length, G=G, dt=p["dt"], engine="binned")
    hp.set_external_fields(None, None)
    hp.initialize(p["x0"], p["v0"])
    out = hp.run(T, particles=False)
    torch.cuda.synchronize()
    for k in ("electric_field", "magnetic_field", "current_density", "charge_density"):
        got = out[k].cpu().numpy()
        scale = max(np.abs(ref[k]).max(), 1e-300)
        if k == "magnetic_field" and scale < 1e-12 * np.abs(ref["electric_field"]).max() / 2.99792458e8:
            assert np.abs(got).max() <= 1e-10 * np.abs(ref["electric_field"]).max() / 2.99792458e8  # no transverse dynamics: B stays (numerically) zero
            continue
        assert np.abs(got - ref[k]).max() / scale < 1e-5, (name, k)
    hp.close()


# ---- step granularity: Boris_step(carry, ...) with the reference's signature (jic_load_carry + one step) ----------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("bcs,relativistic,field_solver,G", [((0, 0, 0, 0), False, 0, 14), ((1, 2, 1, 2), False, 0, 14), ((2, 1, 2, 1), True, 0, 14),
                                                             ((0, 0, 0, 0), False, 2, 14), ((0, 0, 0, 0), False, 0, 1024)])  # 1024: multi-CTA field kernel
def test_boris_step_operator_continues_a_carry_like_the_literal_oracle(bcs, relativistic, field_solver, G):
    """The literal oracle (pinned to the reference's own Boris_step by the refsrc vectors) runs 6 steps and hands over its carry
    (E, B, x_{n-1/2}, x_n, x_{n+1/2}, v, q, m, q/m) -- absorbed particles included; jaxincell_b200.Boris_step continues it for 4 steps, one
    call per step, feeding its own carry back.  step_data and the carry must match the oracle's continuation."""
    from jaxincell_b200 import Boris_step
    from jaxincell_b200._algorithms import release_contexts
    from oracle import literal as L
    from plasma import cfl_dt, two_species
    length = 0.01
    pbl, pbr, fbl, fbr = bcs
    p = two_species(180, 140, length=length, G=G, seed=77 + sum(bcs), vth_e=0.1, vth_yz=0.04, drift=2e7, plus_minus=True, gpdl=0.6)
    dt = cfl_dt(length, G, 0.9)
    solver = dict(filter_passes=3, filter_alpha=0.5, filter_strides=(1, 2), relativistic=relativistic, field_solver=field_solver)
    rng = np.random.default_rng(3)
    ext = {"external_electric_field": (1e3 * rng.standard_normal((G, 3))).astype(np.float32),
           "external_magnetic_field": (1e-3 * rng.standard_normal((G, 3))).astype(np.float32)}
    kw = dict(length=length, G=G, dt=dt, pbl=pbl, pbr=pbr, fbl=fbl, fbr=fbr, solver=solver, ext_E=ext["external_electric_field"],
              ext_B=ext["external_magnetic_field"])
    first = L.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], total_steps=6, **kw)
    full = L.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], total_steps=10, **kw)
    carry = first["final_carry"]
    dx, grid, box = length / G, first["grid"], (length, length, length)
    try:
        for t in range(6, 10):
            carry, (x1, v1, E1, B1, J1, rho1) = Boris_step(carry, t, solver, ext, dx, dt, grid, box, pbl, pbr, fbl, fbr, field_solver)
            for got, key in ((x1, "positions"), (v1, "velocities"), (E1, "electric_field"), (B1, "magnetic_field"), (J1, "current_density"),
                             (rho1, "charge_density")):
                ref = full[key][t]
                assert np.abs(got - ref).max() <= 1e-5 * max(np.abs(ref).max(), 1e-300), (t, key)
        ref_carry = full["final_carry"]
        for got, ref, name in zip(carry, ref_carry, ("E", "B", "x_minus", "x", "x_plus", "v", "q", "m", "qm")):
            got, ref = np.asarray(got, np.float64).reshape(np.shape(ref)), np.asarray(ref, np.float64)
            assert np.abs(got - ref).max() <= 1e-5 * max(np.abs(ref).max(), 1e-300), name
    finally:
        release_contexts()


@pytest.mark.gpu
@pytest.mark.parametrize("bcs", [(0, 0, 0, 0), (2, 2, 2, 2)])
def test_cn_step_operator_continues_a_carry_like_the_literal_oracle(bcs):
    """The same for the implicit stepper: the literal oracle's CN carry after 4 steps, continued for 3 steps by jaxincell_b200.CN_step."""
    from jaxincell_b200 import CN_step
    from jaxincell_b200._algorithms import release_contexts
    from oracle import literal as L
    from plasma import cfl_dt, two_species
    G, length = 16, 0.01
    pbl, pbr, fbl, fbr = bcs
    p = two_species(160, 120, length=length, G=G, seed=90 + sum(bcs), vth_e=0.1 if sum(bcs) else 0.05, vth_yz=0.04, drift=4e7, plus_minus=True, gpdl=0.03)
    dt = cfl_dt(length, G, 0.3)
    solver = {"max_number_of_Picard_iterations_implicit_CN": 12, "number_of_particle_substeps_implicit_CN": 2, "tolerance_Picard_iterations_implicit_CN": 1e-9}
    kw = dict(length=length, G=G, dt=dt, pbl=pbl, pbr=pbr, fbl=fbl, fbr=fbr, solver=solver)
    first = L.run_CN(p["x0"], p["v0"], p["q"], p["m"], p["qm"], total_steps=4, **kw)
    full = L.run_CN(p["x0"], p["v0"], p["q"], p["m"], p["qm"], total_steps=7, **kw)
    carry = first["final_carry"]
    dx, grid, box = length / G, first["grid"], (length, length, length)
    try:
        for t in range(4, 7):
            carry, data = CN_step(carry, t, solver, dx, dt, grid, box, pbl, pbr, fbl, fbr, 2)
            for got, key in zip(data, ("positions", "velocities", "electric_field", "magnetic_field", "current_density", "charge_density")):
                ref = full[key][t]
                assert np.abs(got - ref).max() <= 1e-5 * max(np.abs(ref).max(), 1e-300), (t, key)
    finally:
        release_contexts()


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["indexed", "binned"])
@pytest.mark.parametrize("bcs", [(0, 1, 0, 1), (1, 0, 1, 0), (2, 0, 2, 0), (0, 2, 0, 2)])
def test_field_solver_with_one_periodic_particle_wall(bcs, engine):
    """field_solver != 0 deposits rho(x_n) on the faces; in step 0 x_n is the raw x_0, which the step kernels' recomputation misses for
    particles that the start-up half step sent through a periodic wall opposite a non-periodic one.  k_start_face_fix repairs exactly
    that (found and checked on the CPU emulation of the source: tests/test_cuda_source_on_cpu.py); here on hardware, both engines."""
    import torch
    from jaxincell_b200 import HotPath
    from oracle import closed_form as C
    from plasma import cfl_dt, two_species
    G, length, T = 12, 0.01, 8
    pbl, pbr, fbl, fbr = bcs
    p = two_species(400, 300, length=length, G=G, seed=61 + sum(bcs), vth_e=0.3, vth_yz=0.05, gpdl=0.6)
    dt = cfl_dt(length, G, 0.95)
    solver = dict(field_solver=1, filter_passes=2, filter_strides=(1, 2))
    ref = C.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T, pbl=pbl, pbr=pbr, fbl=fbl, fbr=fbr, solver=solver)
    hp = HotPath(species=p["species"], length=length, G=G, dt=dt, pbl=pbl, pbr=pbr, fbl=fbl, fbr=fbr, engine=engine, track_yz=engine == "indexed",
                 field_solver=1, filter_passes=2, filter_strides=(1, 2))
    hp.set_external_fields(None, None)
    hp.initialize(p["x0"], p["v0"])
    out = hp.run(T, particles=engine == "indexed")
    torch.cuda.synchronize()
    for k in ("electric_field", "magnetic_field", "current_density", "charge_density"):
        ref_k = ref[k]
        assert np.abs(out[k].cpu().numpy() - ref_k).max() <= 1e-5 * max(np.abs(ref_k).max(), 1e-300), (k, engine)
    hp.close()


def _worker_face_fix(rank, world, port, bcs, engine, q):
    """Two ranks on the configuration of the test above: every rank corrects its own particles, the reduced sum stays on rank 0."""
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from jaxincell_b200 import HotPath, shard_particles, shard_species
        from oracle import closed_form as C
        from plasma import cfl_dt, two_species
        G, length, T = 12, 0.01, 8
        pbl, pbr, fbl, fbr = bcs
        p = two_species(400, 300, length=length, G=G, seed=61 + sum(bcs), vth_e=0.3, vth_yz=0.05, gpdl=0.6)
        dt = cfl_dt(length, G, 0.95)
        solver = dict(field_solver=1, filter_passes=2, filter_strides=(1, 2))
        x, v, _ = shard_particles(p["x0"], p["v0"], p["species"], rank, world)
        hp = HotPath(species=shard_species(p["species"], rank, world), length=length, G=G, dt=dt, pbl=pbl, pbr=pbr, fbl=fbl, fbr=fbr, engine=engine,
                     field_solver=1, filter_passes=2, filter_strides=(1, 2))
        hp.comm_init_from_torch()
        hp.set_external_fields(None, None)
        hp.initialize(x, v)
        out = hp.run(T)
        torch.cuda.synchronize()
        if rank == 0:
            ref = C.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T, pbl=pbl, pbr=pbr, fbl=fbl, fbr=fbr,
                        solver=solver, keep_particles=False)
            for k in ("electric_field", "magnetic_field", "current_density", "charge_density"):
                err = np.abs(out[k].cpu().numpy() - ref[k]).max() / max(np.abs(ref[k]).max(), 1e-300)
                assert err < 1e-5, (k, err)
        hp.close()
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, f"{type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.timeout(300)
@pytest.mark.parametrize("engine", ["indexed", "binned"])
@pytest.mark.parametrize("bcs", [(0, 1, 0, 1), (2, 0, 2, 0)])
def test_two_gpus_field_solver_with_one_periodic_particle_wall(bcs, engine):
    import socket
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_face_fix, args=(r, 2, port, bcs, engine, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=240) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["indexed", "binned"])
@pytest.mark.parametrize("seed", [68, 213])
def test_start_up_position_parked_on_a_cell_border(seed, engine):
    """Absorbing left wall: x_{-1/2} = BCpos(x_0 - dt/2 v) of a particle that keeps its charge is parked at grid[0] - 1.5 dx, exactly on a
    border of the J_x window; its cell must be what the reference's float `//` gives (deposit_jx_startup, DevParams::park_left_cell).
    Found by the extended fuzz of the CUDA source on the CPU (tests/test_cuda_source_on_cpu.py, JIC_FUZZ_SCALE=6): the two cases it hit,
    G = 3 at CFL 2.5 with external fields, relativistic."""
    import torch
    from jaxincell_b200 import HotPath
    from oracle import closed_form as C
    from test_cuda_source_on_cpu import _random_case
    g = _random_case(seed)
    pbl, pbr, fbl, fbr = (int(b) for b in g["bcs"])
    strides = tuple(int(s) for s in g["filter_strides"])
    ref = C.run(g["x0"], g["v0"], g["q"], g["m"], g["qm"], length=g["length"], G=g["G"], dt=g["dt"], total_steps=g["T"], pbl=pbl, pbr=pbr, fbl=fbl, fbr=fbr,
                box_yz=tuple(g["box_yz"]), ext_E=g["ext_E"], ext_B=g["ext_B"], keep_particles=False,
                solver=dict(filter_passes=g["filter_passes"], filter_alpha=g["filter_alpha"], filter_strides=strides, relativistic=bool(g["relativistic"])))
    species = [dict(count=g["n_e"], q=float(g["q"][0]), m=float(g["m"][0]), qm=float(g["qm"][0])),
               dict(count=g["n_i"], q=float(g["q"][-1]), m=float(g["m"][-1]), qm=float(g["qm"][-1]))]
    hp = HotPath(species=species, length=g["length"], length_y=float(g["box_yz"][0]), length_z=float(g["box_yz"][1]), G=g["G"], dt=g["dt"], pbl=pbl, pbr=pbr,
                 fbl=fbl, fbr=fbr, engine=engine, track_yz=engine == "indexed", relativistic=bool(g["relativistic"]), filter_passes=g["filter_passes"],
                 filter_alpha=g["filter_alpha"], filter_strides=strides)
    hp.set_external_fields(g["ext_E"], g["ext_B"])
    hp.initialize(g["x0"], g["v0"])
    out = hp.run(g["T"], particles=False)
    torch.cuda.synchronize()
    for k in ("electric_field", "current_density", "charge_density"):
        assert np.abs(out[k].cpu().numpy() - ref[k]).max() <= 1e-5 * max(np.abs(ref[k]).max(), 1e-300), (k, engine)
    hp.close()
