// BINNED particle store (sm_100a): particles live in bins keyed by (species, cell of x_{n+1/2}) and are re-binned by
// the push kernel itself every step, so that for a whole CTA chunk
//   * the gather stencil is the same 4 grid rows  -> E(d), B(d) are quadratics in the in-cell offset d with CTA-uniform
//     coefficients held in registers (no per-particle field loads),
//   * the deposition stencil is the same 5 nodes  -> J_x, J_y, J_z, rho accumulate in per-thread REGISTERS with static
//     indices, are reduced with warp shuffles + shared memory, and reach the L2-resident raw grid as 19 atomics per chunk
//     (instead of ~15 atomics per particle),
//   * species constants (q w, q/m) are uniform,
// and a particle is stored as (d, v_x, v_y, v_z): its cell is implicit, d = (x - g_c)/dx in [-1/2, 1/2].
// Each particle's state crosses HBM once per step (4 reals in, 4 reals out); the out-write goes to the particle's NEW bin
// (slots claimed through warp-aggregated cursor atomics), which is what keeps the store exactly binned with no sort pass.
//
// The arithmetic is that of jaxincell/_algorithms.py:40-66,90-92 (see jic_device.cuh for the per-function citations); the
// "fast path" below is the closed form of the reference's 6-node windowed prefix sum for a particle that moves by at most
// one cell and stays clear of non-periodic walls.  Everything else (multi-cell jumps, wall cells, overflowed bins) takes
// the exact general code of the INDEXED engine, particle by particle -- the deposit is additive, so paths can be mixed.
#pragma once
#include <algorithm>
#include <cstring>
#include <vector>

#include "jic_device.cuh"
#include "jic_host.cuh"
#include "jic_kernels.cuh"

namespace jic {

constexpr int kMinChunk = 1024;    // particles per work item: chosen by k_plan in [kMinChunk, kMaxChunk]
constexpr int kMaxChunk = 8192;
#ifndef JIC_PUSH_THREADS
#define JIC_PUSH_THREADS 64
#endif
#ifndef JIC_PUSH_MINBLOCKS
#define JIC_PUSH_MINBLOCKS 6
#endif
constexpr int kPushThreads = JIC_PUSH_THREADS;
constexpr int kPushMinBlocks = JIC_PUSH_MINBLOCKS;
constexpr int kNumCoef = 21;       // CTA-uniform gather polynomial coefficients (see k_push_binned)

struct PlanHeader {
  int flip;            // which buffer is the SOURCE of the next push
  int n_items;         // work items of the next push
  int ov_n[2];         // entries in the overflow list of each buffer
  int error;           // sticky: 1 = overflow list full, 2 = capacity exhausted
  int chunk;           // particles per work item of the next push
  int work;            // dynamic work queue head (reset by k_plan)
  int pad;
  long long n_stored;  // live particles (bins + overflow list) in the source buffer
  long long n_absorbed;
};

template <typename R>
struct BinDev {
  int nb;                 // n_species * G
  long long cap_total;    // slots per buffer
  int ov_cap;             // overflow list capacity
  float slack;            // head-room fraction per neighbour
  R* d[2]; R* vx[2]; R* vy[2]; R* vz[2];
  long long* off[2];      // [nb+1] first slot of each bin
  int* cnt[2];            // [nb]   particles stored in each bin
  unsigned* cur[2];       // [nb]   write cursors (count every attempt, also the overflowed ones)
  int* ov_bin[2]; R* ov_d[2]; R* ov_vx[2]; R* ov_vy[2]; R* ov_vz[2];
  int* item_bin; int* item_first; int item_cap;
  int n_cta;              // CTAs of the push kernel (work-queue consumers)
  PlanHeader* hdr;
};

__device__ __forceinline__ double rcp_fast(double a) {  // 1/a for a >= 1: MUFU seed + two Newton steps (<= 1 ulp)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  double e = fma(-a, r, 1.0);
  r = fma(r, e, r);
  e = fma(-a, r, 1.0);
  return fma(r, e, r);
}
__device__ __forceinline__ float rcp_fast(float a) { return __frcp_rn(a); }

template <typename R>
__device__ __forceinline__ R node_pos(int c, const DevParams<R>& p) { return p.g0 + R(c) * p.dx; }

// Put one particle into bin `b` of the destination buffer (or into its overflow list when the bin is full).
template <typename R>
__device__ __forceinline__ void store_slot(const BinDev<R>& bd, int dst, int b, unsigned slot, R d, R vx, R vy, R vz) {
  const long long o = bd.off[dst][b];
  const long long cap = bd.off[dst][b + 1] - o;
  if ((long long)slot < cap) {
    const long long k = o + slot;
    bd.d[dst][k] = d; bd.vx[dst][k] = vx; bd.vy[dst][k] = vy; bd.vz[dst][k] = vz;
  } else {
    const int k = atomicAdd(&bd.hdr->ov_n[dst], 1);
    if (k < bd.ov_cap) {
      bd.ov_bin[dst][k] = b; bd.ov_d[dst][k] = d; bd.ov_vx[dst][k] = vx; bd.ov_vy[dst][k] = vy; bd.ov_vz[dst][k] = vz;
    } else {
      atomicExch(&bd.hdr->error, 1);
    }
  }
}

template <typename R>
__device__ __forceinline__ void insert_particle(const BinDev<R>& bd, int dst, int species, R x, R vx, R vy, R vz, const DevParams<R>& p) {
  int c = (int)floor((x - p.gs) * p.inv_dx);
  c = min(max(c, 0), p.G - 1);
  const R d = (x - node_pos(c, p)) * p.inv_dx;
  const int b = species * p.G + c;
  const unsigned slot = atomicAdd(&bd.cur[dst][b], 1u);
  store_slot(bd, dst, b, slot, d, vx, vy, vz);
}

// General (exact) tail of one particle after its velocity update: BC, x_{n+1}, deposit through global atomics, re-insert.
template <typename R>
__device__ __noinline__ void slow_tail(const DevParams<R>& p, const BinDev<R>& bd, int dst, R* acc, int species, R x_old, R v0, R v1, R v2) {
  R v[3] = {v0, v1, v2};
  R x_new = x_old + p.dt * v[0];
  const int flag = bc_x(x_new, p);
  R q = p.sp_q[species];
  if (flag == 1) v[0] = -v[0];
  if (flag == 2) { q = R(0); }
  if (q != R(0)) {
    R x_mid = x_new - p.half_dt * v[0];
    bc_x(x_mid, p);
    const Cloud<R> c_old = make_cloud(x_old, p), c_new = make_cloud(x_new, p), c_mid = make_cloud(x_mid, p);
    const R a = q * p.inv_dx;
    const GlobalGrid<R> g{acc};
    deposit_jx(g, x_old, c_old, c_new, q / p.dt, p);
    deposit_cloud(g, c_mid, p.G, a * v[1], a * v[2], a, true);
    insert_particle(bd, dst, species, x_new, v[0], v[1], v[2], p);
  } else {
    atomicAdd((unsigned long long*)&bd.hdr->n_absorbed, 1ull);  // absorbed: leaves the store, contributes nothing from now on
  }
}

template <typename R>
__device__ __forceinline__ R warp_sum(R v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------------------
// K1b  the binned push: gather -> Boris -> move -> deposit -> re-bin, one pass.
// ---------------------------------------------------------------------------------------------------------
template <typename R, bool REL>
__global__ void __launch_bounds__(kPushThreads, kPushMinBlocks) k_push_binned(const DevParams<R> p, const BinDev<R> bd, const R* __restrict__ F,
                                                                   R* __restrict__ acc) {
  PlanHeader* hdr = bd.hdr;
  const int src = hdr->flip, dst = src ^ 1;
  const R* __restrict__ sd = bd.d[src]; const R* __restrict__ svx = bd.vx[src];
  const R* __restrict__ svy = bd.vy[src]; const R* __restrict__ svz = bd.vz[src];
  const int n_items = hdr->n_items, chunk = hdr->chunk;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int G = p.G;
  const bool periodic = (p.pbl == JIC_BC_PERIODIC) && (p.pbr == JIC_BC_PERIODIC);
  const R cells_per_v = p.dt * p.inv_dx;  // displacement in cells per unit velocity
  __shared__ R red[kPushThreads / 32][20];
  __shared__ R coef[24];
  __shared__ int s_item;
  __shared__ long long s_off[3];
  __shared__ unsigned s_cap[3];

  for (;;) {
    // ---- next work item from the queue (dynamic: items differ in size)
    if (threadIdx.x == 0) s_item = atomicAdd(&hdr->work, 1);
    __syncthreads();
    const int item = s_item;
    if (item >= n_items) break;
    const int b = bd.item_bin[item];
    const int first = bd.item_first[item];
    const int s = b / G, c = b - s * G;
    const int n = min(chunk, bd.cnt[src][b] - first);
    const long long base = bd.off[src][b] + first;
    const bool fast_bin = G >= 8 && (periodic || (c >= 2 && c <= G - 3));
    const int bl = s * G + (c == 0 ? G - 1 : c - 1), br = s * G + (c == G - 1 ? 0 : c + 1);

    // first particle of every thread is already in flight while the coefficients are set up
    R nd = R(0), nv0 = R(0), nv1 = R(0), nv2 = R(0);
    if ((int)threadIdx.x < n) { const long long k = base + threadIdx.x; nd = sd[k]; nv0 = svx[k]; nv1 = svy[k]; nv2 = svz[k]; }

    // ---- CTA-uniform gather polynomials.  Rows c..c+3 of the padded table are f[c-2], f[c-1], f[c], f[c+1].
    //   E lives on faces: for d < 0 the stencil is faces (c-2, c-1, c), for d >= 0 faces (c-1, c, c+1):
    //     E(d) = 1/2 (f[c-1]+f[c]) + d (f[c]-f[c-1]) + d^2 a2,   a2 = 1/2 (f[c-2]+f[c]) - f[c-1]  (d<0),  1/2 (f[c-1]+f[c+1]) - f[c]  (d>=0)
    //   B lives on centres (c-1, c, c+1):
    //     B(d) = 1/8 (b[c-1]+b[c+1]) + 3/4 b[c] + d/2 (b[c+1]-b[c-1]) + d^2 (1/2 (b[c-1]+b[c+1]) - b[c])
    //   coef[k*7 + {0: e0, 1: e1, 2: e2lo, 3: e2hi, 4: b0, 5: b1, 6: b2}], k = component.  Non-relativistic: pre-scaled by (q/m) dt/2.
    if (threadIdx.x < kNumCoef) {
      const int k = threadIdx.x / 7, w_ = threadIdx.x - 7 * k;
      const R hs = REL ? R(1) : p.sp_qm[s] * p.half_dt;
      const R* f = F + (size_t)c * kFieldRow + k;
      R val;
      if (w_ < 4) {
        const R f0 = __ldg(f), f1 = __ldg(f + kFieldRow), f2 = __ldg(f + 2 * kFieldRow), f3 = __ldg(f + 3 * kFieldRow);
        val = w_ == 0 ? R(0.5) * (f1 + f2) : w_ == 1 ? (f2 - f1) : w_ == 2 ? (R(0.5) * (f0 + f2) - f1) : (R(0.5) * (f1 + f3) - f2);
      } else {
        const R b1 = __ldg(f + kFieldRow + 3), b2 = __ldg(f + 2 * kFieldRow + 3), b3 = __ldg(f + 3 * kFieldRow + 3);
        val = w_ == 4 ? (R(0.125) * (b1 + b3) + R(0.75) * b2) : w_ == 5 ? (R(0.5) * (b3 - b1)) : (R(0.5) * (b1 + b3) - b2);
      }
      coef[threadIdx.x] = hs * val;
    } else if (threadIdx.x >= 32 && threadIdx.x < 35) {
      const int k = threadIdx.x - 32, bk = k == 0 ? b : (k == 1 ? bl : br);
      const long long o = bd.off[dst][bk];
      s_off[k] = o;
      s_cap[k] = (unsigned)(bd.off[dst][bk + 1] - o);
    }
    __syncthreads();

    R a_rho[5] = {0, 0, 0, 0, 0}, a_jy[5] = {0, 0, 0, 0, 0}, a_jz[5] = {0, 0, 0, 0, 0}, a_jx[4] = {0, 0, 0, 0};

    // A particle is STORED one iteration after its slot was claimed, so the cursor atomic's round trip overlaps the next
    // particle's arithmetic instead of stalling the warp (q_* = claimed but not yet stored).  The first slot and capacity
    // of the three fast-path destinations (stay / left / right) sit in shared memory (s_off, s_cap), indexed by kind.
    R q_d = R(0), q_v0 = R(0), q_v1 = R(0), q_v2 = R(0);
    int q_kind = -1;
    unsigned q_m = 0, q_c = 0;
    auto retire = [&]() {
      // the leader lane of every destination group holds the base it claimed; one shuffle with a per-lane source
      const unsigned base_slot = __shfl_sync(0xffffffffu, q_c, q_m ? __ffs(q_m) - 1 : 0);
      if (q_kind >= 0) {
        const unsigned slot = base_slot + __popc(q_m & lt_mask);
        if (slot < s_cap[q_kind]) {
          const long long k = s_off[q_kind] + slot;
          bd.d[dst][k] = q_d; bd.vx[dst][k] = q_v0; bd.vy[dst][k] = q_v1; bd.vz[dst][k] = q_v2;
        } else {
          store_slot(bd, dst, q_kind == 0 ? b : (q_kind == 1 ? bl : br), slot, q_d, q_v0, q_v1, q_v2);  // -> overflow list
        }
      }
    };

    const int n_pad = (n + 31) & ~31;
    for (int i = threadIdx.x; i < n_pad; i += kPushThreads) {
      const bool valid = i < n;
      const R d = nd;
      R v[3] = {nv0, nv1, nv2};
      {  // software pipeline: the next particle's loads overlap this particle's arithmetic
        const int j = i + kPushThreads;
        if (j < n) { const long long k = base + j; nd = sd[k]; nv0 = svx[k]; nv1 = svy[k]; nv2 = svz[k]; }
      }
      // ---- gather (quadratics in d)
      R E[3], B[3];
      const bool hi = d >= R(0);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const R a2 = hi ? coef[7 * k + 3] : coef[7 * k + 2];
        E[k] = fma(fma(a2, d, coef[7 * k + 1]), d, coef[7 * k]);
        B[k] = fma(fma(coef[7 * k + 6], d, coef[7 * k + 5]), d, coef[7 * k + 4]);
      }
      // ---- velocity update
      if (REL) {
        boris_velocity_relativistic(v, E, B, p.sp_q[s], p.sp_m[s], p.dt);
      } else {
        // E, B already carry the factor (q/m) dt/2:  v- = v + E ; t = B ; v+ = (R x t + (R.t) t + R)/(1 + t.t) ; v = v+ + E
        const R vm0 = v[0] + E[0], vm1 = v[1] + E[1], vm2 = v[2] + E[2];
        const R R0 = fma(vm1, B[2], fma(-vm2, B[1], vm0)), R1 = fma(vm2, B[0], fma(-vm0, B[2], vm1)), R2 = fma(vm0, B[1], fma(-vm1, B[0], vm2));
        const R Rt = fma(R0, B[0], fma(R1, B[1], R2 * B[2]));
        const R inv = rcp_fast(fma(B[0], B[0], fma(B[1], B[1], fma(B[2], B[2], R(1)))));
        v[0] = fma(fma(R1, B[2], fma(-R2, B[1], fma(Rt, B[0], R0))), inv, E[0]);
        v[1] = fma(fma(R2, B[0], fma(-R0, B[2], fma(Rt, B[1], R1))), inv, E[1]);
        v[2] = fma(fma(R0, B[1], fma(-R1, B[0], fma(Rt, B[2], R2))), inv, E[2]);
      }
      // ---- move (in cell units)
      const R u = v[0] * cells_per_v;
      const R d_new = d + u, d_mid = fma(R(0.5), u, d);
      const bool fast = valid && fast_bin && (fabs(d_new) < R(1.5));
      int kind = -1;  // 0 stay, 1 left, 2 right, 3 general
      if (valid) kind = 3;
      R dn = d_new;
      if (fast) {
        const bool sr_ = d_new >= R(0.5), sl_ = d_new < R(-0.5), mr_ = d_mid >= R(0.5), ml_ = d_mid < R(-0.5);
        const int sh = (int)sr_ - (int)sl_;
        const int shm = (int)mr_ - (int)ml_;
        kind = sr_ ? 2 : (sl_ ? 1 : 0);
        dn = d_new - (sr_ ? R(1) : (sl_ ? R(-1) : R(0)));
        const R dm = d_mid - (mr_ ? R(1) : (ml_ ? R(-1) : R(0)));
        // rho, J_y, J_z: S2 weights of x_{n+1} on nodes c-2..c+2 (its nearest node is c+shm)
        const R wl = R(0.5) * (R(0.5) - dm) * (R(0.5) - dm), wc = R(0.75) - dm * dm, wr = R(0.5) * (R(0.5) + dm) * (R(0.5) + dm);
        const bool ml = shm < 0, mc = shm == 0, mr = shm > 0;
        R w[5];
        w[0] = ml ? wl : R(0);
        w[1] = ml ? wc : (mc ? wl : R(0));
        w[2] = ml ? wr : (mc ? wc : wl);
        w[3] = mc ? wr : (mr ? wc : R(0));
        w[4] = mr ? wr : R(0);
#pragma unroll
        for (int j = 0; j < 5; ++j) { a_rho[j] += w[j]; a_jy[j] = fma(w[j], v[1], a_jy[j]); a_jz[j] = fma(w[j], v[2], a_jz[j]); }
        // J_x on nodes c-2..c+1: difference of the cumulative S2 weights of x_{n+3/2} and x_{n+1/2}
        const R An = R(0.5) * (R(0.5) - dn) * (R(0.5) - dn), Bn = R(1) - R(0.5) * (R(0.5) + dn) * (R(0.5) + dn);
        const R Ao = R(0.5) * (R(0.5) - d) * (R(0.5) - d), Bo = R(1) - R(0.5) * (R(0.5) + d) * (R(0.5) + d);
        const bool sl = sh < 0, sc = sh == 0;
        a_jx[0] += sl ? An : R(0);
        a_jx[1] += (sl ? Bn : (sc ? An : R(0))) - Ao;
        a_jx[2] += (sl ? R(1) : (sc ? Bn : An)) - Bo;
        a_jx[3] += (sl || sc) ? R(0) : (Bn - R(1));
      }
      // ---- claim slots in the destination bins: one atomic per warp and destination (issued by the first lane of each
      //      destination group), consumed next iteration
      const unsigned m0 = __ballot_sync(0xffffffffu, kind == 0), m1 = __ballot_sync(0xffffffffu, kind == 1),
                     m2 = __ballot_sync(0xffffffffu, kind == 2);
      const unsigned m = kind == 0 ? m0 : (kind == 1 ? m1 : (kind == 2 ? m2 : 0u));
      unsigned cl = 0;
      if (m && lane == __ffs(m) - 1) cl = atomicAdd(bd.cur[dst] + (kind == 0 ? b : (kind == 1 ? bl : br)), (unsigned)__popc(m));
      retire();  // the previous particle: its atomic has had a whole iteration to come back
      q_d = dn; q_v0 = v[0]; q_v1 = v[1]; q_v2 = v[2];
      q_kind = kind < 3 ? kind : -1;
      q_m = m; q_c = cl;
      if (kind == 3) slow_tail(p, bd, dst, acc, s, node_pos(c, p) + d * p.dx, v[0], v[1], v[2]);
    }
    retire();

    // ---- flush the register accumulators: warp shuffles -> shared memory -> 19 atomics on the raw grid
    R vals[19];
#pragma unroll
    for (int j = 0; j < 5; ++j) { vals[j] = a_rho[j]; vals[5 + j] = a_jy[j]; vals[10 + j] = a_jz[j]; }
#pragma unroll
    for (int j = 0; j < 4; ++j) vals[15 + j] = a_jx[j];
#pragma unroll
    for (int j = 0; j < 19; ++j) {
      const R t = warp_sum(vals[j]);
      if (lane == 0) red[warp][j] = t;
    }
    __syncthreads();
    if (threadIdx.x < 19) {
      R t = R(0);
#pragma unroll
      for (int w_ = 0; w_ < kPushThreads / 32; ++w_) t += red[w_][threadIdx.x];
      if (t != R(0)) {
        const int j = threadIdx.x;
        const R q = p.sp_q[s];
        int node, comp;
        R scale;
        if (j < 15) { node = c - 2 + (j % 5); comp = j < 5 ? 3 : (j < 10 ? 1 : 2); scale = q * p.inv_dx; }
        else { node = c - 2 + (j - 15); comp = 0; scale = -(q / p.dt); }
        atomicAdd(acc + mod_pos(node, G) * kAccRow + comp, scale * t);
      }
    }
    // (the __syncthreads at the head of the next iteration protects red[] and coef[])
  }

  // ---- particles that did not fit their bin last step: general path, one by one
  const int n_ov = min(hdr->ov_n[src], bd.ov_cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_ov; i += gridDim.x * blockDim.x) {
    const int b = bd.ov_bin[src][i];
    const int s = b / G, c = b - s * G;
    const R x_old = node_pos(c, p) + bd.ov_d[src][i] * p.dx;
    R v[3] = {bd.ov_vx[src][i], bd.ov_vy[src][i], bd.ov_vz[src][i]};
    R E[3], B[3];
    gather_fields(F, x_old, p, E, B);
    if (REL) boris_velocity_relativistic(v, E, B, p.sp_q[s], p.sp_m[s], p.dt);
    else boris_velocity(v, E, B, p.sp_qm[s], p.dt);
    slow_tail(p, bd, dst, acc, s, x_old, v[0], v[1], v[2]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// K3  plan (single CTA): close the buffer that was just written, lay out the NEXT destination buffer with head-room
//     proportional to the population of each bin and its neighbours, build the work-item list, flip.
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__device__ T block_exclusive_scan(T v, T* total, T* smem /* blockDim.x */) {
  // simple Hillis-Steele over blockDim.x partial values (called once or twice per step on 1024 threads)
  const int t = threadIdx.x, n = blockDim.x;
  smem[t] = v;
  __syncthreads();
  for (int o = 1; o < n; o <<= 1) {
    T x = t >= o ? smem[t - o] : T(0);
    __syncthreads();
    smem[t] += x;
    __syncthreads();
  }
  const T incl = smem[t];
  if (total) *total = smem[n - 1];
  __syncthreads();
  return incl - v;
}

template <typename R>
__global__ void __launch_bounds__(1024) k_plan(const BinDev<R> bd, int G, int first_call) {
  __shared__ long long sh_ll[1024];
  __shared__ int sh_i[1024];
  __shared__ long long tot_ll;
  __shared__ int tot_i;
  PlanHeader* h = bd.hdr;
  const int t = threadIdx.x, nt = blockDim.x, nb = bd.nb;
  const int written = first_call ? h->flip : (h->flip ^ 1);  // buffer the last kernel wrote = source of the next push
  const int next = written ^ 1;                               // destination of the next push
  const int per = (nb + nt - 1) / nt, lo = min(t * per, nb), hi = min(lo + per, nb);
  // 1. close `written`: cnt = min(cursor, capacity); everything beyond sits in its overflow list
  long long mine = 0;
  for (int b = lo; b < hi; ++b) {
    const long long cap = bd.off[written][b + 1] - bd.off[written][b];
    const long long att = bd.cur[written][b];
    bd.cnt[written][b] = (int)(att < cap ? att : cap);
    mine += att;
  }
  block_exclusive_scan<long long>(mine, &tot_ll, sh_ll);
  const long long n_total = tot_ll;
  // 2. capacities of `next`: population + slack * (itself and both neighbours in the same species) + a constant
  double f = bd.slack;
  {
    const double room = (double)bd.cap_total - (double)n_total - 40.0 * nb;
    const double fmax = n_total > 0 ? room / (3.0 * (double)n_total) : 0.0;
    if (f > fmax) f = fmax;
    if (f < 0) { f = 0; if (t == 0 && room < 0) atomicExch(&h->error, 2); }
  }
  long long cap_sum = 0;
  for (int b = lo; b < hi; ++b) {
    const int s = b / G, c = b - s * G;
    const long long a0 = bd.cur[written][b];
    const long long al = bd.cur[written][s * G + (c == 0 ? G - 1 : c - 1)], ar = bd.cur[written][s * G + (c == G - 1 ? 0 : c + 1)];
    long long cap = a0 + (long long)(f * (double)(a0 + al + ar)) + 32;
    cap = (cap + 3) & ~3ll;  // keep bins 32-byte aligned
    cap_sum += cap;
  }
  long long run = block_exclusive_scan<long long>(cap_sum, &tot_ll, sh_ll);
  for (int b = lo; b < hi; ++b) {
    const int s = b / G, c = b - s * G;
    const long long a0 = bd.cur[written][b];
    const long long al = bd.cur[written][s * G + (c == 0 ? G - 1 : c - 1)], ar = bd.cur[written][s * G + (c == G - 1 ? 0 : c + 1)];
    long long cap = a0 + (long long)(f * (double)(a0 + al + ar)) + 32;
    cap = (cap + 3) & ~3ll;
    bd.off[next][b] = run;
    run += cap;
  }
  if (t == 0) bd.off[next][nb] = tot_ll;
  __syncthreads();
  for (int b = lo; b < hi; ++b) bd.cur[next][b] = 0u;
  // 3. work items over `written`: about 4 per CTA of the push kernel, between kMinChunk and kMaxChunk particles each
  long long want = n_total / (4ll * (bd.n_cta > 0 ? bd.n_cta : 1));
  want = want < kMinChunk ? kMinChunk : (want > kMaxChunk ? kMaxChunk : want);
  const int kChunk = (int)((want + kPushThreads - 1) / kPushThreads) * kPushThreads;
  int my_items = 0;
  for (int b = lo; b < hi; ++b) my_items += (bd.cnt[written][b] + kChunk - 1) / kChunk;
  int it = block_exclusive_scan<int>(my_items, &tot_i, sh_i);
  for (int b = lo; b < hi; ++b) {
    const int n = bd.cnt[written][b];
    for (int k = 0; k < n; k += kChunk) {
      if (it < bd.item_cap) { bd.item_bin[it] = b; bd.item_first[it] = k; }
      ++it;
    }
  }
  if (t == 0) {
    if (tot_i > bd.item_cap) atomicExch(&h->error, 2);
    h->n_items = tot_i < bd.item_cap ? tot_i : bd.item_cap;
    h->chunk = kChunk;
    h->work = 0;
    h->flip = written;
    h->ov_n[next] = 0;
    h->n_stored = n_total;
  }
}

// ---------------------------------------------------------------------------------------------------------
// start-up: leap-frog start + initial deposits (as k_start) into a linear staging area, then a scatter into bins
// ---------------------------------------------------------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(256) k_start_binned(const DevParams<R> p, const BinDev<R> bd, const R* __restrict__ x0, const R* __restrict__ v0,
                                                      R* __restrict__ st_x, R* __restrict__ st_vx, R* __restrict__ st_vy, R* __restrict__ st_vz,
                                                      int* __restrict__ st_bin, R* __restrict__ acc) {
  const GlobalGrid<R> grid{acc};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.N; i += (long long)gridDim.x * blockDim.x) {
    const int s = species_of(i, p);
    const R q = p.sp_q[s];
    const R X0 = x0[3 * i];
    R v[3] = {v0[3 * i], v0[3 * i + 1], v0[3 * i + 2]};
    const Cloud<R> c0 = make_cloud(X0, p);
    deposit_cloud(grid, c0, p.G, R(0), R(0), q * p.inv_dx, false);
    R xp = X0 + p.half_dt * v[0];
    const int flag = bc_x(xp, p);
    R qj = q;
    if (flag == 1) v[0] = -v[0];
    if (flag == 2) { v[0] = v[1] = v[2] = R(0); qj = R(0); }
    R xm = X0 - p.half_dt * v[0];
    bc_x(xm, p);
    int bin = -1;
    if (qj != R(0)) {
      const Cloud<R> cm = make_cloud(xm, p), cp = make_cloud(xp, p);
      deposit_jx(grid, xm, cm, cp, qj / p.dt, p);
      const R a = qj * p.inv_dx;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int k = c0.c + j - 1;
        if (k >= 0 && k < p.G) { grid.add(k, 1, c0.w[j] * a * v[1]); grid.add(k, 2, c0.w[j] * a * v[2]); }
      }
      if (c0.first != R(0)) { grid.add(0, 1, c0.first * a * v[1]); grid.add(0, 2, c0.first * a * v[2]); }
      if (c0.last != R(0)) { grid.add(p.G - 1, 1, c0.last * a * v[1]); grid.add(p.G - 1, 2, c0.last * a * v[2]); }
      int c = (int)floor((xp - p.gs) * p.inv_dx);
      c = min(max(c, 0), p.G - 1);
      bin = s * p.G + c;
      atomicAdd(&bd.cur[0][bin], 1u);  // histogram of the first layout
    } else {
      atomicAdd((unsigned long long*)&bd.hdr->n_absorbed, 1ull);
    }
    st_x[i] = xp; st_vx[i] = v[0]; st_vy[i] = v[1]; st_vz[i] = v[2]; st_bin[i] = bin;
  }
}

// exact layout for the very first buffer: capacity = histogram count (+ alignment), written by a tiny single-CTA scan
template <typename R>
__global__ void __launch_bounds__(1024) k_first_layout(const BinDev<R> bd) {
  __shared__ long long sh_ll[1024];
  __shared__ long long tot;
  const int t = threadIdx.x, nt = blockDim.x, nb = bd.nb;
  const int per = (nb + nt - 1) / nt, lo = min(t * per, nb), hi = min(lo + per, nb);
  long long mine = 0;
  for (int b = lo; b < hi; ++b) mine += ((long long)bd.cur[0][b] + 3) & ~3ll;
  long long run = block_exclusive_scan<long long>(mine, &tot, sh_ll);
  for (int b = lo; b < hi; ++b) { bd.off[0][b] = run; run += ((long long)bd.cur[0][b] + 3) & ~3ll; }
  if (t == 0) {
    bd.off[0][nb] = tot;
    if (tot > bd.cap_total) atomicExch(&bd.hdr->error, 2);
    bd.hdr->flip = 0;
  }
  __syncthreads();
  for (int b = lo; b < hi; ++b) bd.cur[0][b] = 0u;
}

template <typename R>
__global__ void __launch_bounds__(256) k_scatter_binned(const DevParams<R> p, const BinDev<R> bd, const R* __restrict__ st_x,
                                                        const R* __restrict__ st_vx, const R* __restrict__ st_vy, const R* __restrict__ st_vz,
                                                        const int* __restrict__ st_bin) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.N; i += (long long)gridDim.x * blockDim.x) {
    const int b = st_bin[i];
    if (b < 0) continue;
    const int c = b % p.G;
    const unsigned slot = atomicAdd(&bd.cur[0][b], 1u);
    store_slot(bd, 0, b, slot, (st_x[i] - node_pos(c, p)) * p.inv_dx, st_vx[i], st_vy[i], st_vz[i]);
  }
}

// export / diagnostics over the current source buffer --------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(1024) k_dense_offsets(const BinDev<R> bd, long long* dense /* nb+1 */) {
  __shared__ long long sh_ll[1024];
  __shared__ long long tot;
  const int src = bd.hdr->flip;
  const int t = threadIdx.x, nt = blockDim.x, nb = bd.nb;
  const int per = (nb + nt - 1) / nt, lo = min(t * per, nb), hi = min(lo + per, nb);
  long long mine = 0;
  for (int b = lo; b < hi; ++b) mine += bd.cnt[src][b];
  long long run = block_exclusive_scan<long long>(mine, &tot, sh_ll);
  for (int b = lo; b < hi; ++b) { dense[b] = run; run += bd.cnt[src][b]; }
  if (t == 0) dense[nb] = tot;
}

template <typename R>
__global__ void k_export_binned(const DevParams<R> p, const BinDev<R> bd, const long long* dense, R* x_out, R* v_out, uint8_t* alive) {
  const int src = bd.hdr->flip;
  const long long n_bins = dense[bd.nb];
  for (int b = blockIdx.x; b < bd.nb; b += gridDim.x) {
    const int c = b % p.G;
    const long long o = bd.off[src][b], q0 = dense[b];
    for (int i = threadIdx.x; i < bd.cnt[src][b]; i += blockDim.x) {
      const long long k = q0 + i;
      if (x_out) { x_out[3 * k] = node_pos(c, p) + bd.d[src][o + i] * p.dx; x_out[3 * k + 1] = R(0); x_out[3 * k + 2] = R(0); }
      if (v_out) { v_out[3 * k] = bd.vx[src][o + i]; v_out[3 * k + 1] = bd.vy[src][o + i]; v_out[3 * k + 2] = bd.vz[src][o + i]; }
      if (alive) alive[k] = 1;
    }
  }
  const int n_ov = min(bd.hdr->ov_n[src], bd.ov_cap);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.N - n_bins; i += (long long)gridDim.x * blockDim.x) {
    const long long k = n_bins + i;
    if (i < n_ov) {
      const int c = bd.ov_bin[src][i] % p.G;
      if (x_out) { x_out[3 * k] = node_pos(c, p) + bd.ov_d[src][i] * p.dx; x_out[3 * k + 1] = R(0); x_out[3 * k + 2] = R(0); }
      if (v_out) { v_out[3 * k] = bd.ov_vx[src][i]; v_out[3 * k + 1] = bd.ov_vy[src][i]; v_out[3 * k + 2] = bd.ov_vz[src][i]; }
      if (alive) alive[k] = 1;
    } else {  // absorbed particles have left the store
      if (x_out) { x_out[3 * k] = x_out[3 * k + 1] = x_out[3 * k + 2] = R(0); }
      if (v_out) { v_out[3 * k] = v_out[3 * k + 1] = v_out[3 * k + 2] = R(0); }
      if (alive) alive[k] = 0;
    }
  }
}

template <typename R>
__global__ void k_kinetic_binned(const DevParams<R> p, const BinDev<R> bd, double* out) {
  const int src = bd.hdr->flip;
  double acc = 0.0;
  for (int b = blockIdx.x; b < bd.nb; b += gridDim.x) {
    const double m = (double)p.sp_m[b / p.G];
    const long long o = bd.off[src][b];
    for (int i = threadIdx.x; i < bd.cnt[src][b]; i += blockDim.x) {
      const double a = bd.vx[src][o + i], b_ = bd.vy[src][o + i], c_ = bd.vz[src][o + i];
      acc += 0.5 * m * (a * a + b_ * b_ + c_ * c_);
    }
  }
  const int n_ov = min(bd.hdr->ov_n[src], bd.ov_cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_ov; i += gridDim.x * blockDim.x) {
    const double m = (double)p.sp_m[bd.ov_bin[src][i] / p.G];
    const double a = bd.ov_vx[src][i], b_ = bd.ov_vy[src][i], c_ = bd.ov_vz[src][i];
    acc += 0.5 * m * (a * a + b_ * b_ + c_ * c_);
  }
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

// ---------------------------------------------------------------------------------------------------------
// host side of the store
// ---------------------------------------------------------------------------------------------------------
template <typename R>
struct BinnedStore {
  BinDev<R> bd;
  std::vector<void*> owned;
  long long* dense = nullptr;
  int n_sm = 148;
  bool built = false;

  template <typename T>
  int alloc(Engine& e, T** ptr, size_t n) {
    void* p = nullptr;
    cudaError_t ce = cudaMalloc(&p, (n ? n : 1) * sizeof(T));
    if (ce != cudaSuccess) return e.fail(JIC_ERR_CUDA, format("cudaMalloc(%zu bytes) for the particle bins: %s", n * sizeof(T), cudaGetErrorString(ce)));
    cudaMemset(p, 0, (n ? n : 1) * sizeof(T));
    owned.push_back(p);
    *ptr = (T*)p;
    return JIC_OK;
  }

  int create(Engine& e, const DevParams<R>& dp, const jic_params& prm, int n_sm_) {
    memset(&bd, 0, sizeof(bd));
    n_sm = n_sm_;
    const long long N = dp.N;
    bd.nb = dp.n_species * dp.G;
    bd.slack = 0.125f;
    bd.cap_total = (long long)((double)N * (1.0 + 3.0 * bd.slack)) + 48ll * bd.nb + 1024;
    if (bd.cap_total >= (1ll << 40)) return e.fail(JIC_ERR_UNSUPPORTED, "too many particles for one GPU");
    bd.ov_cap = (int)std::min<long long>(std::max<long long>(N / 16, 1 << 16), 1ll << 28);
    bd.item_cap = (int)std::min<long long>(N / kMinChunk + bd.nb + 16, 1ll << 30);
    bd.n_cta = n_sm * kPushMinBlocks;
    int rc;
    for (int k = 0; k < 2; ++k) {
      if ((rc = alloc(e, &bd.d[k], bd.cap_total)) || (rc = alloc(e, &bd.vx[k], bd.cap_total)) || (rc = alloc(e, &bd.vy[k], bd.cap_total)) ||
          (rc = alloc(e, &bd.vz[k], bd.cap_total)))
        return rc;
      if ((rc = alloc(e, &bd.off[k], bd.nb + 1)) || (rc = alloc(e, &bd.cnt[k], bd.nb)) || (rc = alloc(e, &bd.cur[k], bd.nb))) return rc;
      if ((rc = alloc(e, &bd.ov_bin[k], bd.ov_cap)) || (rc = alloc(e, &bd.ov_d[k], bd.ov_cap)) || (rc = alloc(e, &bd.ov_vx[k], bd.ov_cap)) ||
          (rc = alloc(e, &bd.ov_vy[k], bd.ov_cap)) || (rc = alloc(e, &bd.ov_vz[k], bd.ov_cap)))
        return rc;
    }
    if ((rc = alloc(e, &bd.item_bin, bd.item_cap)) || (rc = alloc(e, &bd.item_first, bd.item_cap)) || (rc = alloc(e, &bd.hdr, 1))) return rc;
    if ((rc = alloc(e, &dense, bd.nb + 1))) return rc;
    (void)prm;
    built = true;
    return JIC_OK;
  }

  void destroy() {
    for (void* p : owned) cudaFree(p);
    owned.clear();
  }

  int grid_for(long long n, int block, int per_sm) const {
    long long b = (n + block - 1) / block, cap = (long long)n_sm * per_sm;
    if (b > cap) b = cap;
    return (int)(b < 1 ? 1 : b);
  }

  // initial binning: staging (in buffer 1, which is free until the first push) -> histogram -> exact layout -> scatter
  int start(Engine& e, const DevParams<R>& dp, const R* x0, const R* v0, R* acc, cudaStream_t st) {
    cudaMemsetAsync(bd.hdr, 0, sizeof(PlanHeader), st);
    cudaMemsetAsync(bd.cur[0], 0, sizeof(unsigned) * bd.nb, st);
    int* st_bin = nullptr;
    cudaError_t ce = cudaMallocAsync((void**)&st_bin, sizeof(int) * (size_t)(dp.N ? dp.N : 1), st);
    if (ce != cudaSuccess) return e.fail(JIC_ERR_CUDA, format("staging allocation: %s", cudaGetErrorString(ce)));
    const int g = grid_for(dp.N, 256, 8);
    k_start_binned<R><<<g, 256, 0, st>>>(dp, bd, x0, v0, bd.d[1], bd.vx[1], bd.vy[1], bd.vz[1], st_bin, acc);
    k_first_layout<R><<<1, 1024, 0, st>>>(bd);
    k_scatter_binned<R><<<g, 256, 0, st>>>(dp, bd, bd.d[1], bd.vx[1], bd.vy[1], bd.vz[1], st_bin);
    cudaFreeAsync(st_bin, st);
    e.launches += 3;
    ce = cudaGetLastError();
    if (ce != cudaSuccess) return e.fail(JIC_ERR_CUDA, format("binned start: %s", cudaGetErrorString(ce)));
    first_plan = true;
    return JIC_OK;
  }
  bool first_plan = true;

  // runs after the field kernel of every step (and of the start-up): plan the next push
  int after_fields(Engine& e, const DevParams<R>& dp, cudaStream_t st) {
    k_plan<R><<<1, 1024, 0, st>>>(bd, dp.G, first_plan ? 1 : 0);
    first_plan = false;
    e.launches += 1;
    return JIC_OK;
  }

  int step(Engine& e, const DevParams<R>& dp, const R* F, R* acc, cudaStream_t st) {
    const int g = bd.n_cta;
    if (dp.relativistic) k_push_binned<R, true><<<g, kPushThreads, 0, st>>>(dp, bd, F, acc);
    else k_push_binned<R, false><<<g, kPushThreads, 0, st>>>(dp, bd, F, acc);
    e.launches += 1;
    return JIC_OK;
  }

  int check_error(Engine& e, cudaStream_t st) {
    PlanHeader h;
    cudaError_t ce = cudaMemcpyAsync(&h, bd.hdr, sizeof(h), cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    if (ce != cudaSuccess) return e.fail(JIC_ERR_CUDA, format("binned store: %s", cudaGetErrorString(ce)));
    if (h.error == 1) return e.fail(JIC_ERR_BAD_STATE, "binned store: overflow list exhausted (a bin grew faster than its head-room)");
    if (h.error == 2) return e.fail(JIC_ERR_BAD_STATE, "binned store: slot capacity exhausted");
    return JIC_OK;
  }

  int export_particles(Engine& e, const DevParams<R>& dp, R* x, R* v, uint8_t* alive, cudaStream_t st) {
    int rc = check_error(e, st);
    if (rc) return rc;
    k_dense_offsets<R><<<1, 1024, 0, st>>>(bd, dense);
    k_export_binned<R><<<n_sm * 4, 256, 0, st>>>(dp, bd, dense, x, v, alive);
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) return e.fail(JIC_ERR_CUDA, format("binned export: %s", cudaGetErrorString(ce)));
    return JIC_OK;
  }

  int kinetic(Engine& e, const DevParams<R>& dp, double* out, cudaStream_t st) {
    k_kinetic_binned<R><<<n_sm * 4, 256, 0, st>>>(dp, bd, out);
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) return e.fail(JIC_ERR_CUDA, format("binned kinetic: %s", cudaGetErrorString(ce)));
    return JIC_OK;
  }

  long long extra_launches_per_step() const { return 1; }  // k_plan
};

}  // namespace jic
