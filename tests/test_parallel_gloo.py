"""N>1 host logic on CPU: world_size-2 `gloo` runs of the sharding scheme of SURVEY.md 8(e).

The multi-GPU path shards every species block by index, deposits locally, all-reduces the RAW [Jx,Jy,Jz,rho] grid and runs the
replicated field solve on the sum.  These tests drive exactly that scheme with the oracle standing in for the per-rank
kernels (no GPU here), through torch.distributed/gloo on 127.0.0.1, and check it against the single-rank oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jaxincell_b200 import shard_counts, shard_particles, shard_species
from jaxincell_b200._parallel import broadcast_bytes
from oracle import closed_form as C
from oracle import literal as L
from plasma import cfl_dt, two_species


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_counts_cover_everything():
    for n in (0, 1, 7, 8, 1001):
        for w in (1, 2, 3, 8):
            c = shard_counts(n, w)
            assert sum(c) == n and max(c) - min(c) <= 1 and len(c) == w


def test_shard_particles_partition_species_blocks():
    p = two_species(11, 6, length=1.0, G=8, seed=3)
    seen = []
    for r in range(3):
        x, v, idx = shard_particles(p["x0"], p["v0"], p["species"], r, 3)
        sp = shard_species(p["species"], r, 3)
        assert len(idx) == sum(s["count"] for s in sp) == len(x) == len(v)
        # rows of one rank are species-contiguous in the same order as the global table
        ne = sp[0]["count"]
        assert (idx[:ne] < 11).all() and (idx[ne:] >= 11).all()
        assert np.array_equal(x, p["x0"][idx])
        seen.append(idx)
    allidx = np.sort(np.concatenate(seen))
    assert np.array_equal(allidx, np.arange(17))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. the courier of the NCCL unique id
        blob = bytes(range(128)) if rank == 0 else None
        got = broadcast_bytes(blob, 128, 0)
        assert got == bytes(range(128))
        # 2. sharded step loop: local push + raw deposit, all-reduce, replicated field solve
        G, length, T = 24, 0.01, 6
        p = two_species(600, 400, length=length, G=G, seed=11, vth_e=0.05, vth_yz=0.02, drift=5e7, plus_minus=True, gpdl=0.7)
        dt = cfl_dt(length, G, 0.8)
        dom = C.Domain(length, G, dt)
        x, v, idx = shard_particles(p["x0"], p["v0"], p["species"], rank, world)
        qs, ms, qms = p["q"][idx], p["m"][idx], p["qm"][idx]
        flt = (5, 0.5, (1, 2, 4), 0, 0)

        def allreduce(a):
            t = torch.from_numpy(np.ascontiguousarray(a))
            dist.all_reduce(t)
            return t.numpy()

        rho0 = L.filter_scalar_field(allreduce(C.deposit_rho_raw(x[:, 0], qs, dom, 0, 0)), *flt)
        E = np.zeros((G, 3)); E[:, 0] = dom.dx / L.epsilon_0 * np.cumsum(rho0)
        B = np.zeros((G, 3))
        xp = L.set_BC_positions(x + dt / 2 * v, dom.dx, dom.grid, *dom.box, 0, 0)
        xm = L.set_BC_positions(x - dt / 2 * v, dom.dx, dom.grid, *dom.box, 0, 0)
        J = L.filter_vector_field(allreduce(C.deposit_current_raw(xm[:, 0], x[:, 0], xp[:, 0], v, qs, dom, 0, 0)), *flt)
        hist = []
        for _ in range(T):
            E, B = L.field_update1(E, B, dom.dx, dt / 2, J, 0, 0)
            Ep, Bp = C.gather_EB(xp[:, 0], E, B, dom, 0, 0)
            xpp, v = C.push_boris(dt, xp, v, qms, Ep, Bp)
            xpp = L.set_BC_positions(xpp, dom.dx, dom.grid, *dom.box, 0, 0)
            xn = L.set_BC_positions(xpp - dt / 2 * v, dom.dx, dom.grid, *dom.box, 0, 0)
            raw = np.concatenate([C.deposit_current_raw(xp[:, 0], xn[:, 0], xpp[:, 0], v, qs, dom, 0, 0),
                                  C.deposit_rho_raw(xn[:, 0], qs, dom, 0, 0)[:, None]], axis=1)  # the (G,4) buffer the kernels all-reduce
            raw = allreduce(raw)
            J = L.filter_vector_field(raw[:, :3], *flt)
            rho = L.filter_scalar_field(raw[:, 3], *flt)
            E, B = L.field_update2(E, B, dom.dx, dt / 2, J, 0, 0)
            hist.append((E.copy(), B.copy(), J.copy(), rho.copy()))
            xp = xpp
        # every rank must hold bit-identical fields (the replicated solve sees the same all-reduced grid)
        mine = torch.from_numpy(np.ascontiguousarray(hist[-1][0]))
        other = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(other, mine)
        assert all(torch.equal(o, mine) for o in other)
        if rank == 0:
            ref = C.run(p["x0"], p["v0"], p["q"], p["m"], p["qm"], length=length, G=G, dt=dt, total_steps=T)
            for t in range(T):
                for k, a in zip(("electric_field", "magnetic_field", "current_density", "charge_density"), hist[t]):
                    scale = max(np.abs(ref[k]).max(), 1e-300)
                    assert np.abs(a - ref[k][t]).max() / scale < 1e-11, (t, k)
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, f"{type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world2_sharded_steps_match_single_rank_oracle():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=240) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
