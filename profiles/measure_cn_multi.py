#!/usr/bin/env python
"""Crank-Nicolson stepper on N GPUs of one box (weak scaling: 2e7 macro-particles per GPU, the workload of measure_rows.py):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 profiles/measure_cn_multi.py
Particles shard by rank; every Picard iteration all-reduces the raw grid over NCCL (the fused peer-memory reduction belongs to the explicit
stepper's field kernel).  Device-timed, max over ranks; rank 0 prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "jax-in-cell_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from jaxincell_b200 import HotPath  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    class A:
        grid, particles = 4096, 20_000_000
    w = bench.workload(A, world)
    x0, v0 = bench.make_particles(w, torch, dev, torch.float64, 1701 + rank, "random")
    S, max_it, steps = 2, 6, 10
    hp = HotPath(species=w["species"], length=w["length"], G=w["G"], dt=0.3 * w["dt"], time_evolution_algorithm=1, cn_substeps=S,
                 cn_max_iterations=max_it, cn_tolerance=1e-30, device=dev)
    if world > 1:
        hp.comm_init_from_torch()
    hp.set_external_fields(None, None)
    hp.initialize(x0, v0)
    del x0, v0
    outs = hp.alloc_outputs(steps)
    hp.run(steps, outputs=outs)
    it0 = hp.picard_iterations()[1]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0.record()
    hp.run(steps, outputs=outs)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    iters = hp.picard_iterations()[1] - it0
    E = outs["electric_field"][-1]
    same = True
    if world > 1:
        others = [torch.empty_like(E) for _ in range(world)]
        dist.all_gather(others, E)
        same = all(torch.equal(o, E) for o in others)
    if rank == 0:
        n_all = hp.N * world
        bytes_per = (12 + 2 * S) * 8 + 1
        t = float(ms.item())
        print(json.dumps({"row": "8f-4 Crank-Nicolson, sorted push, weak scaling", "n_gpus": world, "particles_per_gpu": hp.N, "cn_sorted": hp.store_stats()["cn_sorted"],
                          "substeps": S, "picard_iterations": iters, "ms_per_iteration": t / iters, "particle_iterations_per_s": n_all * iters / t * 1e3,
                          "achieved_GBps_per_gpu": bytes_per * hp.N * iters / t / 1e6, "frac_of_hbm_peak_per_gpu": bytes_per * hp.N * iters / t / 1e6 / bench.peaks()[0],
                          "comm": hp.comm_mode() if world > 1 else "single", "finite": bool(torch.isfinite(E).all()), "ranks_identical": same}), flush=True)
    hp.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
