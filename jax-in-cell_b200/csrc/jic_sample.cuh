// K0s  initial particles on the device, seed-compatible with the reference's sampling.
//
// Restates jaxincell/_state_initialization.py:51-85 (`initialize_species_phase_space`) and the 0.99c clip of :259-260.  The
// random streams are jax.random's, whose arithmetic is NOT in the reference checkout (third-party: jax, unpinned in
// requirements.txt); restated here from its published definition:
//   * PRNGKey(seed) = (seed >> 32, seed & 0xffffffff); the PRNG is Threefry-2x32 with 20 rounds (Salmon et al., SC'11 --
//     pinned by the Random123 known-answer vectors in tests/test_sampling.py);
//   * 64 random bits of element i of a length-n draw: threefry(key; counter) -> (y0 << 32) | y1 with counter (hi32(i), lo32(i))
//     under jax_threefry_partitionable=True (the default since jax 0.5) and counter (i, n + i) under the original layout;
//   * uniform(key, (n,), minval, maxval), float64: u = bitcast((bits >> 12) | 0x3ff0000000000000) - 1, max(minval, u (maxval -
//     minval) + minval);   normal(key, (n,)) = sqrt(2) erfinv(uniform(key, (n,), nextafter(-1, 0), 1));
//   * jnp.linspace(a, b, n)[i] = a (1 - i/(n-1)) + b i/(n-1) for i < n-1, and b exactly at the end.
// erfinv is CUDA's (a few ulp); XLA evaluates its own polynomial, so normal deviates agree to rounding, not bit for bit.
#pragma once
#include "jic_device.cuh"

namespace jic {

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

__device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t& x0, uint32_t& x1) {
  const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
#define JIC_TF_ROUND(r) { x0 += x1; x1 = rotl32(x1, r); x1 ^= x0; }
#define JIC_TF_A JIC_TF_ROUND(13) JIC_TF_ROUND(15) JIC_TF_ROUND(26) JIC_TF_ROUND(6)
#define JIC_TF_B JIC_TF_ROUND(17) JIC_TF_ROUND(29) JIC_TF_ROUND(16) JIC_TF_ROUND(24)
  x0 += k0; x1 += k1;
  JIC_TF_A x0 += k1; x1 += k2 + 1u;
  JIC_TF_B x0 += k2; x1 += k0 + 2u;
  JIC_TF_A x0 += k0; x1 += k1 + 3u;
  JIC_TF_B x0 += k1; x1 += k2 + 4u;
  JIC_TF_A x0 += k2; x1 += k0 + 5u;
#undef JIC_TF_A
#undef JIC_TF_B
#undef JIC_TF_ROUND
}

// jax.random.uniform(PRNGKey(seed), (n,), minval=lo, maxval=hi)[i] in float64
__device__ __forceinline__ double jax_uniform64(long long seed, long long i, long long n, int partitionable, double lo, double hi) {
  const uint32_t k0 = (uint32_t)((unsigned long long)seed >> 32), k1 = (uint32_t)seed;
  uint32_t c0, c1;
  if (partitionable) { c0 = (uint32_t)((unsigned long long)i >> 32); c1 = (uint32_t)i; }
  else { c0 = (uint32_t)i; c1 = (uint32_t)(n + i); }
  threefry2x32(k0, k1, c0, c1);
  const unsigned long long bits = ((unsigned long long)c0 << 32) | c1;
  const double u = __longlong_as_double((long long)((bits >> 12) | 0x3FF0000000000000ull)) - 1.0;
  return fmax(lo, u * (hi - lo) + lo);
}

__device__ __forceinline__ double jax_normal64(long long seed, long long i, long long n, int partitionable) {
  const double lo = -0.99999999999999988897769753748434595763683319091796875;  // nextafter(-1, 0)
  return 1.4142135623730951 * erfinv(jax_uniform64(seed, i, n, partitionable, lo, 1.0));
}

struct SampleArgs {
  long long count, offset;       // particles of this species (the length of the reference's draws), first output row
  long long first, n_local;      // the slice [first, first + n_local) of the species that is generated
  long long seed_position, seed_velocity;
  int random_positions[3], plus_minus[3];
  double amp[3], wavenumber[3];  // perturbation_amplitude, perturbation_wavenumber * 2 pi / box
  double vth[3], drift[3];       // vth_over_c * c / sqrt(2), drift_speed
  double box[3];
  int partitionable;
};

template <typename R>
__global__ void __launch_bounds__(256) k_sample_species(const SampleArgs a, R* __restrict__ x0, R* __restrict__ v0) {
  const double lim = 0.99 * kC;  // _state_initialization.py:259-260
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < a.n_local; j += (long long)gridDim.x * blockDim.x) {
    const long long i = a.first + j;  // index inside the species' draw
    const long long row = 3 * (a.offset + j);
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
      const double half = a.box[ax] / 2;
      double x;
      if (a.random_positions[ax]) {
        x = jax_uniform64(a.seed_position + ax + 1, i, a.count, a.partitionable, -half, half);
      } else if (a.count == 1) {
        x = -half;
      } else if (i == a.count - 1) {
        x = half;
      } else {
        const double t = (double)i / (double)(a.count - 1);
        x = -half * (1.0 - t) + half * t;
      }
      x += a.amp[ax] * sin(a.wavenumber[ax] * x);
      double v = a.vth[ax] * jax_normal64(a.seed_velocity + ax + 4, i, a.count, a.partitionable);
      v += a.drift[ax];
      if (a.plus_minus[ax] && (i & 1)) v = -v;
      if (fabs(v) >= lim) v = v > 0 ? lim : -lim;
      x0[row + ax] = (R)x;
      v0[row + ax] = (R)v;
    }
  }
}

}  // namespace jic
