"""Host-side sharding for N GPUs of one box: particles are split by index inside every species block, the grid is
replicated, and the raw [Jx,Jy,Jz,rho] grid is all-reduced once per step by the library (NCCL).  No reference counterpart:
the reference is single-device."""
from __future__ import annotations

import numpy as np


def shard_counts(n, world):
    """Even split of n items over `world` ranks (first n % world ranks get one more)."""
    base, rem = divmod(int(n), int(world))
    return [base + (1 if r < rem else 0) for r in range(world)]


def shard_species(species, rank, world):
    """Per-rank species table: same charge / mass / q/m (weights are global), local particle counts."""
    return [dict(s, count=shard_counts(s["count"], world)[rank]) for s in species]


def shard_particles(x0, v0, species, rank, world):
    """Rows of the (N,3) arrays owned by `rank`: a contiguous slice of every species block, concatenated block-wise."""
    idx = []
    start = 0
    for s in species:
        counts = shard_counts(s["count"], world)
        lo = start + sum(counts[:rank])
        idx.append(np.arange(lo, lo + counts[rank]))
        start += int(s["count"])
    idx = np.concatenate(idx) if idx else np.zeros(0, dtype=np.int64)
    return np.asarray(x0)[idx], np.asarray(v0)[idx], idx


def broadcast_bytes(payload: bytes | None, nbytes: int, src: int = 0, device=None) -> bytes:
    """Ship a small opaque blob (the 128-byte NCCL unique id) from `src` to every rank through torch.distributed."""
    import torch
    import torch.distributed as dist
    t = torch.zeros(nbytes, dtype=torch.uint8)
    if dist.get_rank() == src:
        t = torch.frombuffer(bytearray(payload), dtype=torch.uint8).clone()
    if dist.get_backend() == "nccl":
        t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    dist.broadcast(t, src=src)
    return t.cpu().numpy().tobytes()
