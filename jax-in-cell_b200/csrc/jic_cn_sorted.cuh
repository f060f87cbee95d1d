// Crank-Nicolson push on CELL-SORTED particles (large runs): the arithmetic of k_cn_push (csrc/jic_cn.cuh; reference
// jaxincell/_algorithms.py:148-188, _sources.py:240-282), with the deposit aggregated per warp.
//
// k_cn_push sends ~21 fp64 atomics per particle and Picard iteration to random nodes of the raw grid: 0.106 of the HBM roofline, bound by
// the L2's atomic rate (profiles/r01_rows_8f_measurements.jsonl).  Here the particle state is re-sorted by cell at the start of every
// step (counting sort: histogram, prefix sum, scatter -- one extra pass over the state per step, against >= 2 Picard iterations that
// each read and write it), so the 32 particles of a warp touch the same handful of nodes.  Every lane adds its contributions to a
// private column of a shared-memory window of kCnWin nodes x 4 components starting three nodes left of the warp's first cell (no
// atomics, no bank conflicts: a column is 33 doubles apart); after the sub-steps each lane sums one (node, component) row over the 32
// columns and issues ONE global atomic: 1 atomic per particle and iteration instead of 21.  A contribution that falls outside the window
// (a particle that wrapped around the periodic box, a parked one, an order gone stale) takes the direct atomic path, so the result is
// exact whatever the order.  Particle histories, exports and the kinetic-energy history go through the permutation (sorted slot ->
// input index) and the per-slot species byte that the sort carries along.
#pragma once
#include "jic_cn.cuh"

namespace jic {

#ifndef JIC_CN_WIN
#define JIC_CN_WIN 6
#endif
#ifndef JIC_CN_SORTED_MINBLOCKS
#define JIC_CN_SORTED_MINBLOCKS 6
#endif
constexpr int kCnWin = JIC_CN_WIN;   // nodes per warp window: a warp whose particles share a cell c touches nodes c-3 .. c+2 (a particle
                                     // moves less than a cell per step: the Picard iteration needs c dt <= dx)
constexpr int kCnWinComps = 4;       // J_x, J_y, J_z, rho
constexpr int kCnColStride = 32;     // a row (node, component) holds one real per lane: the bank of an element depends on the lane alone
#ifndef JIC_CN_PER_LANE
#define JIC_CN_PER_LANE 8
#endif
constexpr int kCnPerLane = JIC_CN_PER_LANE;  // consecutive 32-particle groups a warp runs through one window (zeroed and summed once)
#ifndef JIC_CN_SORTED_THREADS
#define JIC_CN_SORTED_THREADS 128
#endif
constexpr int kCnSortedThreads = JIC_CN_SORTED_THREADS;
static_assert(kCnWin * kCnWinComps <= 32, "at most one (node, component) row per lane in the window reduction");
template <typename R>
__host__ __device__ constexpr size_t cn_sorted_smem_bytes() { return (size_t)(kCnSortedThreads / 32) * kCnWin * kCnWinComps * kCnColStride * sizeof(R); }

#ifndef JIC_CN_PREFETCH
#define JIC_CN_PREFETCH 2
#endif
// hint: bring the line of `ptr` towards the SM (2: L1, 1: L2) for the warp's NEXT chunk -- no register is tied up, unlike a software
// pipeline of loads (the kernel sits at the register limit of its occupancy)
__device__ __forceinline__ void cn_prefetch(const void* ptr) {
#ifdef __CUDA_ARCH__
#if JIC_CN_PREFETCH == 2
  asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr));
#elif JIC_CN_PREFETCH == 1
  asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
#endif
#else
  (void)ptr;
#endif
}

template <typename R>
__device__ __forceinline__ int cn_cell(R x, const DevParams<R>& p) {
  int c = (int)floor((x - p.gs) * p.inv_dx);
  return min(max(c, 0), p.G - 1);
}

// sorted slot -> input index, species of the slot: identity / species blocks before the first sort
template <typename R>
__global__ void __launch_bounds__(256) k_cn_meta_init(const DevParams<R> p, int* __restrict__ perm, uint8_t* __restrict__ sp) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.N; i += (long long)gridDim.x * blockDim.x) {
    perm[i] = (int)i;
    sp[i] = (uint8_t)species_of(i, p);
  }
}

// counting sort by cell of x_n, pass 1: histogram (lanes of a warp that share a cell add once; lanes past the end carry cell -1)
template <typename R>
__global__ void __launch_bounds__(256) k_cn_hist(const DevParams<R> p, const R* __restrict__ x, unsigned* __restrict__ hist) {
  const int lane = threadIdx.x & 31;
  for (long long i0 = blockIdx.x * (long long)blockDim.x; i0 < p.N; i0 += (long long)gridDim.x * blockDim.x) {
    const long long i = i0 + threadIdx.x;
    const int c = i < p.N ? cn_cell(x[i], p) : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, c);
    if (c >= 0 && lane == __ffs(peers) - 1) atomicAdd(hist + c, (unsigned)__popc(peers));
  }
}

// pass 2: exclusive prefix sum of the histogram (single CTA), cursors zeroed
__global__ void __launch_bounds__(1024) k_cn_scan(int G, unsigned* __restrict__ hist, unsigned* __restrict__ off) {
  __shared__ unsigned warp_tot[33];
  __shared__ unsigned carry;
  if (threadIdx.x == 0) carry = 0u;
  __syncthreads();
  for (int base = 0; base < G; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const unsigned v = i < G ? hist[i] : 0u;
    unsigned incl = v;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    if (w == 0) {
      unsigned t = lane < (int)(blockDim.x >> 5) ? warp_tot[lane] : 0u, ti = t;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned u = __shfl_up_sync(0xffffffffu, ti, o);
        if (lane >= o) ti += u;
      }
      warp_tot[lane] = ti - t;
      if (lane == 31) warp_tot[32] = ti;
    }
    __syncthreads();
    if (i < G) { off[i] = carry + warp_tot[w] + incl - v; hist[i] = 0u; }
    __syncthreads();
    if (threadIdx.x == 0) carry += warp_tot[32];
    __syncthreads();
  }
}

// pass 3: scatter the state, the permutation, the species and alive bytes to their cell's run
template <typename R>
__global__ void __launch_bounds__(256) k_cn_scatter(const DevParams<R> p, CnState<R> in, CnState<R> out, const int* __restrict__ perm_in,
                                                    int* __restrict__ perm_out, const uint8_t* __restrict__ sp_in, uint8_t* __restrict__ sp_out,
                                                    const uint8_t* __restrict__ alive_in, uint8_t* __restrict__ alive_out,
                                                    const unsigned* __restrict__ off, unsigned* __restrict__ cursor) {
  const int lane = threadIdx.x & 31;
  for (long long i0 = blockIdx.x * (long long)blockDim.x; i0 < p.N; i0 += (long long)gridDim.x * blockDim.x) {
    const long long i = i0 + threadIdx.x;
    const bool have = i < p.N;
    const R x = have ? in.x[i] : R(0);
    const int c = have ? cn_cell(x, p) : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, c);
    const int leader = __ffs(peers) - 1;
    unsigned base = 0;
    if (have && lane == leader) base = atomicAdd(cursor + c, (unsigned)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (have) {
      const long long k = (long long)off[c] + base + __popc(peers & ((1u << lane) - 1u));
      out.x[k] = x; out.y[k] = in.y[i]; out.z[k] = in.z[i];
      out.vx[k] = in.vx[i]; out.vy[k] = in.vy[i]; out.vz[k] = in.vz[i];
      perm_out[k] = perm_in[i]; sp_out[k] = sp_in[i]; alive_out[k] = alive_in[i];
    }
  }
}

// the stencil of cn_stencil with the quotient taken as a product with 1/dx and the wrap as a compare (the integer remainder and the
// fp64 division cost more than the rest of the gather).  The S2 weights are continuous in x, also across the rounding tie where the
// two quotients may pick different centre nodes, so the deposit and the gather differ from cn_stencil's by rounding only.
template <typename R>
__device__ __forceinline__ int cn_stencil_fast(R x, R start, const DevParams<R>& p, int idx[3], R w[3]) {
  const R xn = (x - start) * p.inv_dx;
  const R kf = rint(xn);
  const int k = (int)kf;
  const R d = xn - kf;
  if (k >= 1 && k + 1 < p.G) { idx[0] = k - 1; idx[1] = k; idx[2] = k + 1; }
  else { idx[0] = mod_pos(k - 1, p.G); idx[1] = mod_pos(k, p.G); idx[2] = mod_pos(k + 1, p.G); }
  w[0] = R(0.5) * (R(0.5) - d) * (R(0.5) - d);
  w[1] = R(0.75) - d * d;
  w[2] = R(0.5) * (R(0.5) + d) * (R(0.5) + d);
  return k;
}

template <typename R> struct CnPair;
template <> struct CnPair<double> { using type = double2; };
template <> struct CnPair<float> { using type = float2; };

// One Picard iteration on sorted particles (see the header).  Same per-particle arithmetic as k_cn_push.
template <typename R>
__global__ void __launch_bounds__(kCnSortedThreads, JIC_CN_SORTED_MINBLOCKS) k_cn_push_sorted(const DevParams<R> p, CnState<R> cur, CnState<R> nxt, R* __restrict__ stag, int n_sub,
                                                                        int it, const double* __restrict__ EB, R* __restrict__ acc, const uint8_t* __restrict__ alive,
                                                                        const uint8_t* __restrict__ sp_of, const CnControl* __restrict__ cn) {
  if (it > 0 && cn->converged) return;
  using R2 = typename CnPair<R>::type;
  extern __shared__ __align__(256) unsigned char cn_sorted_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kRows = kCnWin * kCnWinComps;
  R* win = reinterpret_cast<R*>(cn_sorted_smem) + (size_t)warp * kRows * kCnColStride;  // rows (node, comp) x 32 lanes
  R2* win2 = reinterpret_cast<R2*>(win);
  const R dtau = p.dt / R(n_sub), half_dtau = R(0.5) * dtau;
  const R e_start = p.g0 + p.half_dx;  // _algorithms.py:110 (the B faces, :111, are the E faces shifted by one node: table EB)
  const R w_sub = dtau / p.dt;
  const GlobalGrid<R> grid{acc};
  const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5), warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  constexpr long long kChunk = 32ll * kCnPerLane;  // particles per warp and window pass
  for (long long i0 = warp_global * kChunk; i0 < p.N; i0 += warps_total * kChunk) {
    // zero the window (the warp together, 16 bytes per lane and store)
#pragma unroll
    for (int t = 0; t < kRows * kCnColStride / 64; ++t) win2[t * 32 + lane] = R2{R(0), R(0)};
    int base = 0;  // first node of the window: three left of the lowest cell of the chunk's first 32 particles (the order is ascending;
                   // whatever falls outside the window -- stale order, sparse cells -- takes the direct path)
    // add w3[j] * (a0, a1, a2) to J (with_j) or w3[j] * a0 to rho on nodes k-1, k, k+1 (UNWRAPPED index k; idx = the wrapped ones)
    auto deposit3 = [&](int k, const R w3[3], const int idx[3], R a0, R a1, R a2, bool with_j) {
      const int rel0 = k - 1 - base;
      if (rel0 >= 0 && rel0 + 2 < kCnWin) {
        R* row = win + (rel0 * kCnWinComps) * kCnColStride + lane;
#pragma unroll
        for (int j = 0; j < 3; ++j, row += kCnWinComps * kCnColStride) {
          if (with_j) {
            row[0] += w3[j] * a0; row[kCnColStride] += w3[j] * a1; row[2 * kCnColStride] += w3[j] * a2;
          } else {
            row[3 * kCnColStride] += w3[j] * a0;
          }
        }
      } else {  // outside the warp's window: straight to the grid
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (with_j) {
            if (w3[j] * a0 != R(0)) atomicAdd(acc + idx[j] * kAccRow + 0, w3[j] * a0);
            if (w3[j] * a1 != R(0)) atomicAdd(acc + idx[j] * kAccRow + 1, w3[j] * a1);
            if (w3[j] * a2 != R(0)) atomicAdd(acc + idx[j] * kAccRow + 2, w3[j] * a2);
          } else if (w3[j] * a0 != R(0)) {
            atomicAdd(acc + idx[j] * kAccRow + 3, w3[j] * a0);
          }
        }
      }
    };
    for (int h = 0; h < kCnPerLane; ++h) {
    if (i0 + 32ll * h >= p.N) break;  // (warp-uniform)
    const long long i = i0 + 32ll * h + lane;
    const bool have = i < p.N;
    R pos[3] = {R(0), R(0), R(0)}, vel[3] = {R(0), R(0), R(0)};
    int c0 = 0x3fffffff;
    R xs = R(0);
    if (have) {
      pos[0] = cur.x[i]; pos[1] = cur.y[i]; pos[2] = cur.z[i];
      vel[0] = cur.vx[i]; vel[1] = cur.vy[i]; vel[2] = cur.vz[i];
      xs = it == 0 ? pos[0] : stag[i];
      c0 = cn_cell(pos[0], p);
    }
    if (h == 0) {
      base = __reduce_min_sync(0xffffffffu, c0) - 3;
      __syncwarp();  // (the window is zero for every lane from here on)
    }
    {  // what this warp reads next -- the chunk's next 32 particles, or the first 32 of its next chunk -- on its way meanwhile
      const long long in = h + 1 < kCnPerLane ? i + 32 : i0 + warps_total * kChunk + lane;
      if (in < p.N) {
        cn_prefetch(cur.x + in); cn_prefetch(cur.y + in); cn_prefetch(cur.z + in);
        cn_prefetch(cur.vx + in); cn_prefetch(cur.vy + in); cn_prefetch(cur.vz + in);
        if (it != 0) for (int s = 0; s < n_sub; ++s) cn_prefetch(stag + (size_t)s * p.N + in);
        if ((lane & 7) == 0) { cn_prefetch(sp_of + in); cn_prefetch(alive + in); }
      }
    }
    if (have) {
      const int sp = sp_of[i];
      const bool live = alive[i] != 0;
      const R q = live ? p.sp_q[sp] : R(0), qm = live ? p.sp_qm[sp] : R(0);
      const R a = q * p.inv_dx * w_sub;
      const R x_n = pos[0];
      for (int s = 0; s < n_sub; ++s) {
        // (the staggered position of the NEXT sub-step is fetched before this one's is replaced: the store below would otherwise
        // fence the load behind it)
        const R xs_next = (it != 0 && s + 1 < n_sub) ? stag[(size_t)(s + 1) * p.N + i] : x_n;
        int ie[3];
        R we[3];
        const int ke = cn_stencil_fast(xs, e_start, p, ie, we);  // unwrapped centre node of the face stencil
        R E[3] = {0, 0, 0}, B[3] = {0, 0, 0};
#pragma unroll
        for (int k = 0; k < 3; ++k) {  // row ie[k] of the packed table: E on face ie[k], B on the B face of the same stencil slot
          const double2* row = reinterpret_cast<const double2*>(EB + ie[k] * 6);
          const double2 r0 = __ldg(row), r1 = __ldg(row + 1), r2 = __ldg(row + 2);
          E[0] += we[k] * (R)r0.x; E[1] += we[k] * (R)r0.y; E[2] += we[k] * (R)r1.x;
          B[0] += we[k] * (R)r1.y; B[1] += we[k] * (R)r2.x; B[2] += we[k] * (R)r2.y;
        }
        R vnew[3] = {vel[0], vel[1], vel[2]};
        boris_velocity(vnew, E, B, qm, dtau);
        R vmid[3] = {R(0.5) * (vel[0] + vnew[0]), R(0.5) * (vel[1] + vnew[1]), R(0.5) * (vel[2] + vnew[2])};
        pos[0] += vmid[0] * dtau; pos[1] += vmid[1] * dtau; pos[2] += vmid[2] * dtau;
        const int flag = bc_x(pos[0], p);
        pos[1] = wrap_transverse(pos[1], p.Ly, p.half_Ly);
        pos[2] = wrap_transverse(pos[2], p.Lz, p.half_Lz);
        if (flag == 1) vmid[0] = -vmid[0];
        if (flag == 2) { vmid[0] = vmid[1] = vmid[2] = R(0); }
        R xst = pos[0] - half_dtau * vmid[0];
        bc_x(xst, p);
        stag[(size_t)s * p.N + i] = xst;
        deposit3(ke, we, ie, a * vmid[0], a * vmid[1], a * vmid[2], true);
        vel[0] = vnew[0]; vel[1] = vnew[1]; vel[2] = vnew[2];
        xs = xs_next;  // (x_n in the first iteration, _algorithms.py:121)
      }
      nxt.x[i] = pos[0]; nxt.y[i] = pos[1]; nxt.z[i] = pos[2];
      nxt.vx[i] = vel[0]; nxt.vy[i] = vel[1]; nxt.vz[i] = vel[2];
      // rho(x_{n+1}): the S2 cloud on the centres with the particle-BC fold (as k_cn_push); its three interior nodes go through the
      // window, the folded ghost weights (wall cells only) straight to the grid
      const Cloud<R> cl = make_cloud(pos[0], p);
      const R ar = q * p.inv_dx;
      bool interior = cl.c - 1 >= 0 && cl.c + 1 < p.G;
      if (interior && cl.first == R(0) && cl.last == R(0)) {
        const int idx[3] = {cl.c - 1, cl.c, cl.c + 1};
        deposit3(cl.c, cl.w, idx, ar, R(0), R(0), false);
      } else {
        deposit_cloud(grid, cl, p.G, R(0), R(0), ar, false);
      }
    }
    }  // h
    __syncwarp();
    // lane l sums row l (node l / 4, component l % 4) over the 32 columns, two at a time; the XOR with the lane spreads the lanes of
    // a quarter warp over the eight 16-byte bank groups
    if (lane < kRows) {
      const R2* row = win2 + lane * (kCnColStride / 2);
      R s0 = R(0), s1 = R(0);
#pragma unroll
      for (int u = 0; u < kCnColStride / 2; ++u) {
        const R2 v = row[u ^ (lane & 15)];
        s0 += v.x; s1 += v.y;
      }
      const R sum = s0 + s1;
      if (sum != R(0)) atomicAdd(acc + mod_pos(base + (lane >> 2), p.G) * kAccRow + (lane & 3), sum);
    }
    __syncwarp();
  }
}

// histories in INPUT order through the permutation
template <typename R>
__global__ void __launch_bounds__(256) k_cn_record_sorted(const DevParams<R> p, CnState<R> s, const int* __restrict__ perm, const RunControl* __restrict__ ctl) {
  R* x_hist = (R*)ctl->hist[4];
  R* v_hist = (R*)ctl->hist[5];
  if (!x_hist && !v_hist) return;
  const long long row = ctl->hist_row - 1;  // k_cn_fields has already advanced it
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.N; i += (long long)gridDim.x * blockDim.x) {
    const long long k = perm[i];
    if (x_hist) { R* o = x_hist + ((size_t)row * p.N + k) * 3; o[0] = s.x[i]; o[1] = s.y[i]; o[2] = s.z[i]; }
    if (v_hist) { R* o = v_hist + ((size_t)row * p.N + k) * 3; o[0] = s.vx[i]; o[1] = s.vy[i]; o[2] = s.vz[i]; }
  }
}

template <typename R>
__global__ void __launch_bounds__(256) k_cn_export_sorted(const DevParams<R> p, CnState<R> s, const int* __restrict__ perm, R* x_out, R* v_out, uint8_t* alive) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.N; i += (long long)gridDim.x * blockDim.x) {
    const long long k = perm[i];
    const R x = s.x[i];
    if (x_out) { x_out[3 * k] = x; x_out[3 * k + 1] = s.y[i]; x_out[3 * k + 2] = s.z[i]; }
    if (v_out) { v_out[3 * k] = s.vx[i]; v_out[3 * k + 1] = s.vy[i]; v_out[3 * k + 2] = s.vz[i]; }
    if (alive) alive[k] = !((x < -p.half_L) || (x > p.half_L));
  }
}

// 0.5 sum m v^2 with the species byte of the slot: per species into `out_hist` row (kinetic-energy history) or all into out[0]
template <typename R>
__global__ void __launch_bounds__(256) k_cn_kinetic_sorted(const DevParams<R> p, const R* vx, const R* vy, const R* vz, const uint8_t* __restrict__ sp_of,
                                                           double* total, const RunControl* ctl, int row_back) {
  double* hist = ctl ? (double*)ctl->hist[6] : nullptr;
  if (!total && !hist) return;
  if (hist) hist += (ctl->hist_row - row_back) * p.n_species;
  double acc[JIC_MAX_SPECIES];
#pragma unroll
  for (int s = 0; s < JIC_MAX_SPECIES; ++s) acc[s] = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.N; i += (long long)gridDim.x * blockDim.x) {
    const int s = sp_of[i];
    const double a = vx[i], b = vy[i], c = vz[i];
    const double e = 0.5 * (double)p.sp_m[s] * (a * a + b * b + c * c);
#pragma unroll
    for (int k = 0; k < JIC_MAX_SPECIES; ++k) acc[k] += k == s ? e : 0.0;
  }
#pragma unroll
  for (int s = 0; s < JIC_MAX_SPECIES; ++s) {
    if (s < p.n_species) {
      double v = acc[s];
      for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(hist ? hist + s : total, v);
    }
  }
}

}  // namespace jic
