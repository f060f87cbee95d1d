"""2, 4 or 8 ranks, one GPU each, one reduction of the raw grid per step (NCCL all-reduce, or fused into the field kernel over peer
memory): fields must match the single-rank oracle and be bit-identical across ranks.  Every world size the box has the
GPUs for runs; the others are skipped (gpurun --gpus 2 / 4 / 8)."""
import os
import socket

import numpy as np
import pytest

from oracle import closed_form as C
from plasma import cfl_dt, two_species

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, engine, q, G=64, p2p=True):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), JIC_P2P="1" if p2p else "0")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from jaxincell_b200 import HotPath, shard_particles, shard_species
        length, T = 0.01, 12
        p = two_species(6000, 5000, length=length, G=G, seed=21, vth_e=0.05, vth_yz=0.02, drift=5e7, plus_minus=True, gpdl=0.7)
        dt = cfl_dt(length, G, 0.9)
        x, v, idx = shard_particles(p["x0"], p["v0"], p["species"], rank, world)
        hp = HotPath(species=shard_species(p["species"], rank, world), length=length, G=G, dt=dt, engine=engine)
        hp.comm_init_from_torch()
        # grids large enough for the multi-CTA field kernel reduce inside it over peer memory; small ones go through NCCL
        mode = hp.comm_mode()
        assert mode == ("fused" if (p2p and G >= 256) else "nccl"), mode
        hp.set_external_fields(None, None)
        hp.initialize(x, v)
        a = hp.run(5)
        hp.initialize(x, v)          # a second run on the same context (what bench.py's e2e leg does)
        a = hp.run(5)
        b = hp.run(T - 5)
        out = {k: torch.cat([a[k], b[k]]) for k in a}
        torch.cuda.synchronize()
        E = out["electric_field"]
        others = [torch.empty_like(E) for _ in range(world)]
        dist.all_gather(others, E)
        assert all(torch.equal(o, E) for o in others), "ranks hold different fields"
        if rank == 0:
            ref = C.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T, keep_particles=False)
            for k in ("electric_field", "magnetic_field", "current_density", "charge_density"):
                err = np.abs(out[k].cpu().numpy() - ref[k]).max() / max(np.abs(ref[k]).max(), 1e-300)
                assert err < 1e-5, (k, err)  # north-star tolerance, fp64
        hp.close()
        q.put((rank, "ok:" + mode))
    except Exception as e:  # noqa: BLE001
        q.put((rank, f"{type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("engine,G,p2p", [("indexed", 64, True), ("binned", 64, True), ("binned", 512, True), ("indexed", 300, True),
                                          ("binned", 512, False), ("binned", 4096, True)])
def test_n_gpus_match_single_rank_oracle(engine, G, p2p, world):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    if world > 2 and (engine, G) not in (("binned", 512), ("binned", 4096)):
        pytest.skip("the small-grid / INDEXED variants are covered at two ranks")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, engine, q, G, p2p)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=240) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
    want = "ok:fused" if (p2p and G >= 256) else "ok:nccl"
    assert sorted(res) == [(r, want) for r in range(world)], res


def _worker_modes(rank, world, port, mode, q):
    """field_solver and the Crank-Nicolson stepper on two ranks (both reduce through NCCL)."""
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from jaxincell_b200 import HotPath, shard_particles, shard_species
        from oracle import literal as L
        G, length, T = 300, 0.01, 8
        p = two_species(6000, 5000, length=length, G=G, seed=21, vth_e=0.05, vth_yz=0.02, drift=5e7, plus_minus=True, gpdl=0.05)
        dt = cfl_dt(length, G, 0.3)
        x, v, idx = shard_particles(p["x0"], p["v0"], p["species"], rank, world)
        kw = dict(field_solver=1, engine="binned") if mode == "field_solver" else dict(time_evolution_algorithm=1, cn_max_iterations=6, cn_tolerance=1e-8)
        if mode == "crank_nicolson_sorted":  # the large-run push (csrc/jic_cn_sorted.cuh) on every rank's shard, whatever its size
            os.environ["JIC_CN_SORTED_MIN"] = "0"
        hp = HotPath(species=shard_species(p["species"], rank, world), length=length, G=G, dt=dt, **kw)
        hp.comm_init_from_torch()
        assert hp.comm_mode() == "nccl"
        if mode != "field_solver":
            assert hp.store_stats()["cn_sorted"] == (1 if mode == "crank_nicolson_sorted" else 0)
        hp.set_external_fields(None, None)
        hp.initialize(x, v)
        out = hp.run(T)
        torch.cuda.synchronize()
        E = out["electric_field"]
        others = [torch.empty_like(E) for _ in range(world)]
        dist.all_gather(others, E)
        assert all(torch.equal(o, E) for o in others), "ranks hold different fields"
        if rank == 0:
            if mode == "field_solver":
                ref = C.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T, keep_particles=False,
                            solver=dict(field_solver=1))
            else:
                ref = L.run_CN(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T,
                               solver=dict(max_number_of_Picard_iterations_implicit_CN=6, tolerance_Picard_iterations_implicit_CN=1e-8))
            for k in ("electric_field", "magnetic_field", "current_density", "charge_density"):
                err = np.abs(out[k].cpu().numpy() - ref[k]).max() / max(np.abs(ref[k]).max(), 1e-300)
                assert err < 1e-5, (k, err)
        hp.close()
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, f"{type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("mode", ["field_solver", "crank_nicolson", "crank_nicolson_sorted"])
def test_n_gpus_field_solver_and_crank_nicolson(mode, world):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_modes, args=(r, world, port, mode, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=240) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res


def _worker_bump(rank, world, port, n_total, T, q):
    """BASELINE config 4 as BASELINE.json states it: the bump-on-tail plasma sharded by index over the ranks."""
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import bump_on_tail as BT
        from jaxincell_b200 import HotPath, shard_species
        from jaxincell_b200._parallel import shard_counts
        s = BT.setup(n_total)
        # every rank draws its own slice on its device (one seed per rank: the plasma is statistically the same as the single-GPU one)
        local = shard_species(s["species"], rank, world)
        s_local = dict(s, counts=tuple(sp["count"] for sp in local))
        dev = torch.device("cuda", rank)
        x0, v0 = BT.particles_torch(s_local, dev, seed=250724 + rank)
        hp = HotPath(engine="binned", species=local, length=s["length"], G=s["G"], dt=s["dt"], filter_passes=0)
        hp.comm_init_from_torch()
        hp.set_external_fields(None, None)
        hp.initialize(x0, v0)
        del x0, v0
        out = hp.run(T)
        hp.check_status()
        E = out["electric_field"]
        ref0 = E.clone()
        dist.broadcast(ref0, src=0)
        same = torch.tensor([1 if torch.equal(ref0, E) else 0], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        rho_sum = float(out["charge_density"].double().sum(dim=1).abs().max()) * s["dx"]
        q_scale = sum(sp["count"] * abs(sp["q"]) for sp in s["species"])
        gamma, mode, window = BT.growth_rate_of_mode(E[:, :, 0].cpu().numpy(), s)
        st = hp.store_stats()
        hp.close()
        q.put((rank, dict(identical=bool(int(same.item())), neutrality=rho_sum / q_scale, gamma=gamma / s["gamma_theory"], mode=mode,
                          theory_mode=s["mode"], window=window, error=st["error"], comm=world)))
    except Exception as e:  # noqa: BLE001
        q.put((rank, f"{type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world", [2, 4, 8])
def test_config4_bump_on_tail_sharded(world):
    """2e8 macro-particles over 2 / 4 / 8 GPUs (BASELINE config 4): ranks hold bit-identical fields, the plasma stays neutral, and the
    k = 4.9 omega_pe / c mode grows at the rate of linear theory (0.075 omega_pe, tolerance as in tests/test_full_size.py)."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    sys_path_tests = os.path.dirname(os.path.abspath(__file__))
    os.environ["PYTHONPATH"] = sys_path_tests + os.pathsep + os.environ.get("PYTHONPATH", "")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_bump, args=(r, world, port, 200_000_000, 320, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = sorted((q.get(timeout=540) for _ in procs), key=lambda t: t[0])
    for pr in procs:
        pr.join(timeout=60)
    for rank, r in res:
        assert isinstance(r, dict), r
        assert r["identical"] and r["error"] == 0, r
        assert r["neutrality"] < 1e-9, r
        assert abs(r["mode"] - r["theory_mode"]) <= 1, r
        assert abs(r["gamma"] - 1) < 0.15, r
    print("bump-on-tail sharded:", res[0][1])
