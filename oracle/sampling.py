"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference's initial particle sampling.

Only ``tests/`` may import this.  Follows jaxincell/_state_initialization.py:51-96 (`initialize_species_phase_space`,
`species_seed_pair`) and :259-260 (0.99c clip).  The random streams come from ``jax.random`` -- a third-party dependency that is
NOT under /root/reference (``jax``, unpinned in requirements.txt) and cannot be installed here.  Its published algorithm is
restated: Threefry-2x32-20 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11; the Random123
known-answer vectors pin it), jax's key/counter layout, bits -> float conversion and normal = sqrt(2) erfinv(uniform(-1, 1)).

Parity status: threefry itself PINNED (Random123 KATs); the scalar float32 draws quoted in the JAX documentation
(uniform(PRNGKey(0)) = 0.41845703 with the original bit layout, 0.947667 with jax_threefry_partitionable; normal(PRNGKey(0)) =
-0.20584226) are reproduced by `uniform32_scalar` / `normal32_scalar` below; the float64 vector layout follows the same code path
of jax/_src/prng.py but has no vector from the reference to check against -> "parity unpinned" for the composed sampler.
"""
from __future__ import annotations

import numpy as np
from scipy.special import erfinv

speed_of_light = 2.99792458e8

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32, 20 rounds.  Scalars or uint32 arrays; returns (y0, y1) as uint32 arrays."""
    with np.errstate(over="ignore"):
        k0 = np.uint32(k0); k1 = np.uint32(k1)
        x0 = np.array(x0, dtype=np.uint32, copy=True); x1 = np.array(x1, dtype=np.uint32, copy=True)
        ks = (k0, k1, np.uint32(k0 ^ k1 ^ np.uint32(0x1BD11BDA)))
        x0 = x0 + ks[0]; x1 = x1 + ks[1]
        for g in range(5):
            for r in _ROT[g % 2]:
                x0 = x0 + x1
                x1 = (x1 << np.uint32(r)) | (x1 >> np.uint32(32 - r))
                x1 = x1 ^ x0
            x0 = x0 + ks[(g + 1) % 3]
            x1 = x1 + ks[(g + 2) % 3] + np.uint32(g + 1)
    return x0, x1


def _key(seed):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.uint32(seed >> 32), np.uint32(seed & 0xFFFFFFFF)


def random_bits64(seed, n, partitionable=True):
    k0, k1 = _key(seed)
    i = np.arange(n, dtype=np.uint64)
    if partitionable:
        c0, c1 = (i >> np.uint64(32)).astype(np.uint32), (i & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    else:
        c0, c1 = i.astype(np.uint32), (i + np.uint64(n)).astype(np.uint32)
    y0, y1 = threefry2x32(k0, k1, c0, c1)
    return (y0.astype(np.uint64) << np.uint64(32)) | y1.astype(np.uint64)


def uniform64(seed, n, lo, hi, partitionable=True):
    bits = random_bits64(seed, n, partitionable)
    u = ((bits >> np.uint64(12)) | np.uint64(0x3FF0000000000000)).view(np.float64) - 1.0
    return np.maximum(lo, u * (hi - lo) + lo)


def normal64(seed, n, partitionable=True):
    lo = np.nextafter(-1.0, 0.0)
    return np.sqrt(2.0) * erfinv(uniform64(seed, n, lo, 1.0, partitionable))


def uniform32_scalar(seed, partitionable):
    """jax.random.uniform(PRNGKey(seed)) with the default float32 dtype and shape (): one 32-bit word."""
    k0, k1 = _key(seed)
    y0, y1 = threefry2x32(k0, k1, np.uint32(0), np.uint32(0))
    bits = np.uint32(y0 ^ y1) if partitionable else np.uint32(y0)
    return float(np.array((bits >> np.uint32(9)) | np.uint32(0x3F800000), dtype=np.uint32).view(np.float32) - np.float32(1.0))


def normal32_scalar(seed, partitionable):
    lo = np.nextafter(np.float32(-1), np.float32(0))
    u = np.float32(uniform32_scalar(seed, partitionable))
    u = max(lo, np.float32(u * (np.float32(1) - lo) + lo))
    return float(np.float32(np.sqrt(2)) * np.float32(erfinv(np.float64(u))))


def jnp_linspace(a, b, n):
    if n == 1:
        return np.array([a], dtype=np.float64)
    t = np.arange(n - 1, dtype=np.float64) / np.float64(n - 1)
    return np.concatenate([a * (1.0 - t) + b * t, [b]])


def species_seed_pair(seed, species_type, rng_index, extra_rng_index=None):
    """_state_initialization.py:87-96"""
    if species_type == "electrons" and rng_index == 0:
        return seed, seed + 3
    if species_type == "ions" and rng_index == 0:
        return seed, seed + 6
    if extra_rng_index is None:
        extra_rng_index = max(rng_index - 1, 0)
    local = seed + 12 + extra_rng_index * 6
    return local, local


def species_phase_space(sp, box, partitionable=True):
    """One species: (count,3) positions and velocities.  `sp` as for jaxincell_b200.sample_particles."""
    n = int(sp["count"])
    x = np.empty((n, 3)); v = np.empty((n, 3))
    for a in range(3):
        half = box[a] / 2
        if sp["random_positions"][a]:
            xa = uniform64(sp["seed_position"] + a + 1, n, -half, half, partitionable)
        else:
            xa = jnp_linspace(-half, half, n)
        k = sp["perturbation_wavenumber"][a] * 2 * np.pi / box[a]
        xa = xa + sp["perturbation_amplitude"][a] * np.sin(k * xa)
        va = sp["vth_over_c"][a] * speed_of_light / np.sqrt(2) * normal64(sp["seed_velocity"] + a + 4, n, partitionable)
        va = va + sp["drift_speed"][a]
        if sp["velocity_plus_minus"][a]:
            va = va * (-1.0) ** np.arange(n)
        x[:, a], v[:, a] = xa, va
    return x, v


def sample(species, box, partitionable=True):
    xs, vs = zip(*(species_phase_space(sp, box, partitionable) for sp in species))
    x, v = np.concatenate(xs), np.concatenate(vs)
    lim = 0.99 * speed_of_light
    return x, np.where(np.abs(v) >= lim, np.sign(v) * lim, v)
