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
// Memory layout: slots are grouped in BLOCKS of 32 (one warp); a block is [d x32][v_x x32][v_y x32][v_z x32], i.e. structure
// of arrays inside 1 KiB (fp64) records.  Bins start on block boundaries.  A warp therefore reads a bin as a contiguous
// stream of whole blocks (one 1-D bulk async copy per pipeline stage, any length), every lane's four values sit at
// immediate offsets 0/256/512/768 B from one address, and re-binned particles of one warp land in runs of consecutive
// slots that fill whole 32-byte sectors.
//
// The arithmetic is that of jaxincell/_algorithms.py:40-66,90-92 (see jic_device.cuh for the per-function citations); the
// "fast path" below is the closed form of the reference's 6-node windowed prefix sum for a particle that moves by at most
// one cell and stays clear of non-periodic walls.  Everything else (multi-cell jumps, wall cells, overflowed bins) takes
// the exact general code of the INDEXED engine, particle by particle -- the deposit is additive, so paths can be mixed.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "jic_device.cuh"
#include "jic_host.cuh"
#include "jic_kernels.cuh"

namespace jic {

constexpr int kMinChunk = 1024;    // particles per work item: chosen by k_plan in [kMinChunk, kMaxChunk]
#ifndef JIC_MAX_CHUNK
#define JIC_MAX_CHUNK 8192
#endif
constexpr int kMaxChunk = JIC_MAX_CHUNK;
constexpr int kChunkAlign = 256;   // items start on multiples of this inside a bin (a multiple of the block size)
constexpr int kSlowChunk = 256;    // work-item size in bins whose particles all take the general path (wall cells, the edge cells
                                   // of field_solver runs): small, so that this per-particle work spreads over many warps
constexpr int kBlk = 32;           // slots per block
constexpr int kBlkElems = 4 * kBlk;  // reals per block
#ifndef JIC_PUSH_THREADS
#define JIC_PUSH_THREADS 128
#endif
#ifndef JIC_PUSH_MINBLOCKS
#define JIC_PUSH_MINBLOCKS 3
#endif
#ifndef JIC_PUSH_STAGES
#define JIC_PUSH_STAGES 4          // ring slots per warp (power of two)
#endif
#ifndef JIC_PUSH_STAGE_BLOCKS
#define JIC_PUSH_STAGE_BLOCKS 2    // 32-particle blocks per ring slot
#endif
constexpr int kPushThreads = JIC_PUSH_THREADS;
constexpr int kPushMinBlocks = JIC_PUSH_MINBLOCKS;
#ifndef JIC_PUSH_MINBLOCKS_F32
#define JIC_PUSH_MINBLOCKS_F32 5   // the fp32 kernel fits 96 registers: 20 warps per SM (measured best of 3..6)
#endif
template <typename R>
constexpr int push_min_blocks() { return sizeof(R) == 8 ? kPushMinBlocks : JIC_PUSH_MINBLOCKS_F32; }
constexpr int kPushStages = JIC_PUSH_STAGES;
constexpr int kPushStageBlocks = JIC_PUSH_STAGE_BLOCKS;
constexpr int kPushWarps = kPushThreads / 32;

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

constexpr int kPlanMaxCtas = 16;
struct PlanSync {           // k_plan_mc: per-CTA totals and the two counters of its one grid-wide barrier
  long long cap_total[kPlanMaxCtas];
  int item_total[kPlanMaxCtas];
  unsigned arrive, depart;
};

template <typename R>
struct BinDev {
  int nb;                 // n_species * G
  long long cap_total;    // slots per buffer
  int ov_cap;             // overflow list capacity
  float slack;            // head-room fraction per neighbour
  R* rec[2];              // blocked particle records of each buffer: 4 * cap_total reals
  long long* off[2];      // [nb+1] first slot of each bin
  int* cnt[2];            // [nb]   particles stored in each bin
  unsigned* cur[2];       // [nb]   write cursors (count every attempt, also the overflowed ones)
  int* ov_bin[2]; R* ov_d[2]; R* ov_vx[2]; R* ov_vy[2]; R* ov_vz[2];
  int* item_bin; int* item_first; int item_cap;
  int n_workers;          // warps of the push kernel (work-queue consumers)
  struct PlanSync* psync; // cross-CTA state of k_plan_mc
  int edge;               // cells c < edge or c > G-1-edge take the general path for every particle (0 = none, G = all)
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

// component a of slot k lives at slot_ptr(rec, k)[kBlk * a]
template <typename R>
__device__ __forceinline__ R* slot_ptr(R* rec, long long k) { return rec + ((k >> 5) << 7) + (k & 31); }

// Put one particle into bin `b` of the destination buffer (or into its overflow list when the bin is full).
template <typename R>
__device__ __forceinline__ void store_slot(const BinDev<R>& bd, int dst, int b, unsigned slot, R d, R vx, R vy, R vz) {
  const long long o = bd.off[dst][b];
  const long long cap = bd.off[dst][b + 1] - o;
  if ((long long)slot < cap) {
    R* q = slot_ptr(bd.rec[dst], o + slot);
    q[0] = d; q[kBlk] = vx; q[2 * kBlk] = vy; q[3 * kBlk] = vz;
  } else {
    const int k = atomicAdd(&bd.hdr->ov_n[dst], 1);
    if (k < bd.ov_cap) {
      bd.ov_bin[dst][k] = b; bd.ov_d[dst][k] = d; bd.ov_vx[dst][k] = vx; bd.ov_vy[dst][k] = vy; bd.ov_vz[dst][k] = vz;
    } else {
      atomicExch(&bd.hdr->error, 1);
    }
  }
}

// Called from divergent code: the lanes that are here together and head for the same bin claim their slots with ONE cursor
// atomic (a wall cell sends thousands of particles through this path into two or three bins).
template <typename R>
__device__ __forceinline__ void insert_particle(const BinDev<R>& bd, int dst, int species, R x, R vx, R vy, R vz, const DevParams<R>& p) {
  int c = (int)floor((x - p.gs) * p.inv_dx);
  c = min(max(c, 0), p.G - 1);
  const R d = (x - node_pos(c, p)) * p.inv_dx;
  const int b = species * p.G + c;
  const unsigned active = __activemask();
  const unsigned peers = __match_any_sync(active, b);
  const int leader = __ffs(peers) - 1;
  int lane;
  asm("mov.u32 %0, %%laneid;" : "=r"(lane));
  unsigned base = 0;
  if (lane == leader) base = atomicAdd(&bd.cur[dst][b], (unsigned)__popc(peers));
  base = __shfl_sync(peers, base, leader);
  const unsigned slot = base + __popc(peers & ((1u << lane) - 1u));
  store_slot(bd, dst, b, slot, d, vx, vy, vz);
}

// General (exact) tail of one particle after its velocity update: BC, x_{n+1}, deposit through global atomics, re-insert.
template <typename R>
__device__ __noinline__ void slow_tail(const DevParams<R>& p, const BinDev<R>& bd, int dst, R* acc, int species, R x_old, R vx_old, R v0, R v1, R v2) {
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
    const GlobalGrid<R> g{acc, p.G};
    deposit_jx(g, x_old, c_old, c_new, q / p.dt, p);
    deposit_cloud(g, c_mid, p.G, a * v[1], a * v[2], a, true);
    if (p.stag) {  // rho(x_n) on the faces (field_solver != 0), as in k_step
      R x_n = x_old - p.half_dt * vx_old;
      bc_x(x_n, p);
      deposit_faces(g, make_cloud_faces(x_n, p), p.G, a);
    }
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

}  // namespace jic
#include "jic_push.cuh"
namespace jic {

// ---------------------------------------------------------------------------------------------------------
// K3  plan (single CTA): close the buffer that was just written, lay out the NEXT destination buffer with head-room
//     proportional to the population of each bin and its neighbours, build the work-item list, flip.
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__device__ T block_exclusive_scan(T v, T* total, T* smem /* >= 33 */) {
  // warp shuffles inside the warps, one more warp scan over the warp totals (blockDim.x <= 1024, a multiple of 32)
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  T incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const T x = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += x;
  }
  if (lane == 31) smem[w] = incl;
  __syncthreads();
  if (w == 0) {
    const T t = lane < nw ? smem[lane] : T(0);
    T ti = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const T x = __shfl_up_sync(0xffffffffu, ti, o);
      if (lane >= o) ti += x;
    }
    smem[lane] = ti - t;
    if (lane == 31) smem[32] = ti;
  }
  __syncthreads();
  const T res = incl - v + smem[w];
  if (total && threadIdx.x == 0) *total = smem[32];
  __syncthreads();
  return res;
}

// SMEM: the per-bin cursor values and counts are staged in shared memory (2 * nb ints), so that global memory is read once
// and the three phases are not separated by L2 round trips; without it (nb too large) the same code re-reads global memory.
template <typename R, bool SMEM>
__global__ void __launch_bounds__(1024) k_plan(const BinDev<R> bd, int G, int first_call) {
  extern __shared__ int plan_sm[];  // SMEM: [nb] attempts, [nb] counts
  __shared__ long long sh_ll[33];
  __shared__ int sh_i[33];
  __shared__ long long tot_ll;
  __shared__ int tot_i;
  PlanHeader* h = bd.hdr;
  const int t = threadIdx.x, nt = blockDim.x, nb = bd.nb;
  const int written = first_call ? h->flip : (h->flip ^ 1);  // buffer the last kernel wrote = source of the next push
  const int next = written ^ 1;                               // destination of the next push
  const unsigned* __restrict__ g_att = bd.cur[written];
  int* s_att = plan_sm;
  int* s_cnt = plan_sm + nb;
  auto att_of = [&](int b) -> long long { return SMEM ? (long long)(unsigned)s_att[b] : (long long)g_att[b]; };
  auto cnt_of = [&](int b) -> int { return SMEM ? s_cnt[b] : bd.cnt[written][b]; };
  // 1. close `written`: cnt = min(cursor, capacity); everything beyond sits in its overflow list (bins interleaved over threads)
  long long mine = 0;
  {
    // all loads of a thread's first kBatch bins are issued before the first store (the stores would otherwise fence the loads of
    // the next iteration: one L2 round trip per bin on the step's critical path)
    constexpr int kBatch = 8;
    const long long* __restrict__ g_off = bd.off[written];
    long long o0[kBatch], o1[kBatch];
    unsigned at[kBatch];
#pragma unroll
    for (int k = 0; k < kBatch; ++k) {
      const int b = t + k * nt;
      o0[k] = o1[k] = 0; at[k] = 0u;
      if (b < nb) { o0[k] = g_off[b]; o1[k] = g_off[b + 1]; at[k] = g_att[b]; }
    }
#pragma unroll
    for (int k = 0; k < kBatch; ++k) {
      const int b = t + k * nt;
      if (b < nb) {
        const long long cap = o1[k] - o0[k], att = at[k];
        const int cnt = (int)(att < cap ? att : cap);
        bd.cnt[written][b] = cnt;
        if (SMEM) { s_att[b] = (int)att; s_cnt[b] = cnt; }
        mine += att;
      }
    }
    for (int b = t + kBatch * nt; b < nb; b += nt) {
      const long long cap = g_off[b + 1] - g_off[b];
      const long long att = g_att[b];
      const int cnt = (int)(att < cap ? att : cap);
      bd.cnt[written][b] = cnt;
      if (SMEM) { s_att[b] = (int)att; s_cnt[b] = cnt; }
      mine += att;
    }
  }
  block_exclusive_scan<long long>(mine, &tot_ll, sh_ll);
  const long long n_total = tot_ll;
  // 2. capacities of `next`: population + slack * (itself and both neighbours in the same species) + a constant.
  //    From here on every thread owns a contiguous range of bins (offsets are a running sum).
  const int per = (nb + nt - 1) / nt, lo = min(t * per, nb), hi = min(lo + per, nb);
  double f = bd.slack;
  {
    const double room = (double)bd.cap_total - (double)n_total - 72.0 * nb;
    const double fmax = n_total > 0 ? room / (3.0 * (double)n_total) : 0.0;
    if (f > fmax) f = fmax;
    if (f < 0) { f = 0; if (t == 0 && room < 0) atomicExch(&h->error, 2); }
  }
  // (one integer division per thread: the cell index runs along with the bin index from here on; this kernel is a single CTA on
  //  the step's critical path and was instruction-bound -- 275 instructions per bin, mostly divisions)
  const int s_lo = lo / G, c_lo = lo - s_lo * G;
  constexpr int kKeep = 8;  // capacities kept in registers between the two passes (per <= 8 up to 8192 bins)
  long long caps[kKeep];
  auto cap_at = [&](int b, int c) -> long long {
    const long long a0 = att_of(b), al = att_of(c == 0 ? b + G - 1 : b - 1), ar = att_of(c == G - 1 ? b - (G - 1) : b + 1);
    const long long cap = a0 + (long long)(f * (double)(a0 + al + ar)) + 32;
    return (cap + kBlk - 1) & ~(long long)(kBlk - 1);  // bins start on block boundaries
  };
  long long cap_sum = 0;
  {
    int c = c_lo;
#pragma unroll
    for (int k = 0; k < kKeep; ++k) {
      const int b = lo + k;
      caps[k] = 0;
      if (b < hi) { caps[k] = cap_at(b, c); cap_sum += caps[k]; }
      c = c + 1 == G ? 0 : c + 1;
    }
    for (int b = lo + kKeep; b < hi; ++b) { cap_sum += cap_at(b, c); c = c + 1 == G ? 0 : c + 1; }
  }
  long long run = block_exclusive_scan<long long>(cap_sum, &tot_ll, sh_ll);
  {
    int c = c_lo;
#pragma unroll
    for (int k = 0; k < kKeep; ++k) {
      const int b = lo + k;
      if (b < hi) { bd.off[next][b] = run; run += caps[k]; bd.cur[next][b] = 0u; }
      c = c + 1 == G ? 0 : c + 1;
    }
    for (int b = lo + kKeep; b < hi; ++b) { bd.off[next][b] = run; run += cap_at(b, c); bd.cur[next][b] = 0u; c = c + 1 == G ? 0 : c + 1; }
  }
  if (t == 0) bd.off[next][nb] = tot_ll;
  // 3. work items over `written`: about 4 per warp of the push kernel, between kMinChunk and kMaxChunk particles each
  long long want = n_total / (4ll * (bd.n_workers > 0 ? bd.n_workers : 1));
  want = want < kMinChunk ? kMinChunk : (want > kMaxChunk ? kMaxChunk : want);
  const int kChunk = (int)((want + kChunkAlign - 1) / kChunkAlign) * kChunkAlign;
  const float inv_chunk = 1.0f / (float)kChunk;
  static_assert((kSlowChunk & (kSlowChunk - 1)) == 0, "kSlowChunk must be a power of two");
  auto items_of = [&](int n, int c) -> int {  // ceil(n / chunk of this bin) without an integer division
    if (c < bd.edge || c > G - 1 - bd.edge) return (n + kSlowChunk - 1) / kSlowChunk;
    int q = (int)((float)n * inv_chunk);                        // within one of the quotient ...
    while ((long long)q * kChunk < n) ++q;                      // ... made exact
    while (q > 0 && (long long)(q - 1) * kChunk >= n) --q;
    return q;
  };
  int my_items = 0;
  {
    int c = c_lo;
    for (int b = lo; b < hi; ++b) { my_items += items_of(cnt_of(b), c); c = c + 1 == G ? 0 : c + 1; }
  }
  int it = block_exclusive_scan<int>(my_items, &tot_i, sh_i);
  {
    int c = c_lo;
    for (int b = lo; b < hi; ++b) {
      const int n = cnt_of(b), ch = (c < bd.edge || c > G - 1 - bd.edge) ? kSlowChunk : kChunk;
      for (int k = 0; k < n; k += ch) {
        if (it < bd.item_cap) { bd.item_bin[it] = b; bd.item_first[it] = k; }
        ++it;
      }
      c = c + 1 == G ? 0 : c + 1;
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
// K3m  the plan on NC CTAs (same result as k_plan).  The single-CTA version is latency-bound (38 us at 8192 bins) and sits on the
//      step's critical path next to an 18 us field kernel; here every thread owns `per` contiguous bins (one at 8192 bins),
//      neighbour populations come straight from the cursors in global memory, every CTA sums all cursors itself (n_total and
//      the slack factor need no exchange), and the two prefix sums (slot offsets, work items) are two-level: block scan, per-CTA
//      totals through global memory, ONE grid-wide barrier (all CTAs are co-resident: the GPU is otherwise idle at this point).
// ---------------------------------------------------------------------------------------------------------
constexpr int kPlanMcThreads = 512;

template <typename R>
__global__ void __launch_bounds__(kPlanMcThreads) k_plan_mc(const BinDev<R> bd, int G, int first_call) {
  __shared__ long long sh_ll[33];
  __shared__ int sh_i[33];
  __shared__ long long tot_ll;
  __shared__ int tot_i;
  PlanHeader* h = bd.hdr;
  PlanSync* ps = bd.psync;
  const int t = threadIdx.x, nt = blockDim.x, nb = bd.nb, NC = gridDim.x, cta = blockIdx.x;
  const int written = first_call ? h->flip : (h->flip ^ 1);
  const int next = written ^ 1;
  const unsigned* __restrict__ g_att = bd.cur[written];
  const long long* __restrict__ g_off = bd.off[written];
  // n_total: every CTA sums all cursors (a few KB from L2)
  long long mine = 0;
  for (int b = t; b < nb; b += nt) mine += (long long)g_att[b];
  block_exclusive_scan<long long>(mine, &tot_ll, sh_ll);
  const long long n_total = tot_ll;
  double f = bd.slack;
  {
    const double room = (double)bd.cap_total - (double)n_total - 72.0 * nb;
    const double fmax = n_total > 0 ? room / (3.0 * (double)n_total) : 0.0;
    if (f > fmax) f = fmax;
    if (f < 0) { f = 0; if (t == 0 && cta == 0 && room < 0) atomicExch(&h->error, 2); }
  }
  long long want = n_total / (4ll * (bd.n_workers > 0 ? bd.n_workers : 1));
  want = want < kMinChunk ? kMinChunk : (want > kMaxChunk ? kMaxChunk : want);
  const int kChunk = (int)((want + kChunkAlign - 1) / kChunkAlign) * kChunkAlign;
  const float inv_chunk = 1.0f / (float)kChunk;
  // own bins: [lo, hi), contiguous over the whole grid of threads
  const int T = NC * nt, per = (nb + T - 1) / T;
  const int lo = min((cta * nt + t) * per, nb), hi = min(lo + per, nb);
  const int s_lo = lo / G, c_lo = lo - s_lo * G;
  auto is_slow = [&](int c) { return c < bd.edge || c > G - 1 - bd.edge; };
  auto cap_at = [&](int b, int c) -> long long {
    const long long a0 = g_att[b], al = g_att[c == 0 ? b + G - 1 : b - 1], ar = g_att[c == G - 1 ? b - (G - 1) : b + 1];
    const long long cap = a0 + (long long)(f * (double)(a0 + al + ar)) + 32;
    return (cap + kBlk - 1) & ~(long long)(kBlk - 1);
  };
  auto cnt_at = [&](int b) -> int {
    const long long cap = g_off[b + 1] - g_off[b], att = g_att[b];
    return (int)(att < cap ? att : cap);
  };
  auto items_of = [&](int n, int c) -> int {
    if (is_slow(c)) return (n + kSlowChunk - 1) / kSlowChunk;
    int q = (int)((float)n * inv_chunk);
    while ((long long)q * kChunk < n) ++q;
    while (q > 0 && (long long)(q - 1) * kChunk >= n) --q;
    return q;
  };
  long long cap_sum = 0;
  int item_sum = 0;
  {
    int c = c_lo;
    for (int b = lo; b < hi; ++b) {
      const int cnt = cnt_at(b);
      bd.cnt[written][b] = cnt;
      cap_sum += cap_at(b, c);
      item_sum += items_of(cnt, c);
      c = c + 1 == G ? 0 : c + 1;
    }
  }
  long long run = block_exclusive_scan<long long>(cap_sum, &tot_ll, sh_ll);
  int it = block_exclusive_scan<int>(item_sum, &tot_i, sh_i);
  // ---- the one grid-wide exchange: per-CTA totals
  if (t == 0) {
    ps->cap_total[cta] = tot_ll;
    ps->item_total[cta] = tot_i;
    __threadfence();
    atomicAdd(&ps->arrive, 1u);
    const long long t0 = clock64();
    while (*(volatile unsigned*)&ps->arrive < (unsigned)NC) {
      if (clock64() - t0 > 2000000000ll) { atomicExch(&h->error, 2); break; }  // never hang the GPU on a lost CTA
    }
    __threadfence();
  }
  __syncthreads();
  long long all_caps = 0;
  int all_items = 0;
  for (int k = 0; k < NC; ++k) {
    const long long ck = *(volatile long long*)&ps->cap_total[k];
    const int ik = *(volatile int*)&ps->item_total[k];
    if (k < cta) { run += ck; it += ik; }
    all_caps += ck; all_items += ik;
  }
  {
    int c = c_lo;
    for (int b = lo; b < hi; ++b) {
      bd.off[next][b] = run;
      run += cap_at(b, c);
      bd.cur[next][b] = 0u;
      const int n = bd.cnt[written][b], ch = is_slow(c) ? kSlowChunk : kChunk;
      for (int k = 0; k < n; k += ch) {
        if (it < bd.item_cap) { bd.item_bin[it] = b; bd.item_first[it] = k; }
        ++it;
      }
      c = c + 1 == G ? 0 : c + 1;
    }
  }
  if (cta == 0 && t == 0) {
    bd.off[next][nb] = all_caps;
    if (all_items > bd.item_cap) atomicExch(&h->error, 2);
    h->n_items = all_items < bd.item_cap ? all_items : bd.item_cap;
    h->chunk = kChunk;
    h->work = 0;
    h->flip = written;
    h->ov_n[next] = 0;
    h->n_stored = n_total;
  }
  // every CTA has read the totals: the last one to leave re-arms the barrier for the next launch
  __syncthreads();
  if (t == 0) {
    __threadfence();
    if (atomicAdd(&ps->depart, 1u) == (unsigned)NC - 1) { ps->arrive = 0u; ps->depart = 0u; __threadfence(); }
  }
}

// ---------------------------------------------------------------------------------------------------------
// start-up: leap-frog start + initial deposits (as k_start) into a linear staging area, then a scatter into bins
// ---------------------------------------------------------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(256) k_start_binned(const DevParams<R> p, const BinDev<R> bd, const R* __restrict__ x0, const R* __restrict__ v0,
                                                      long long i0, long long n,  // x0, v0 hold particles [i0, i0 + n) of the run
                                                      R* __restrict__ st_x, R* __restrict__ st_vx, R* __restrict__ st_vy, R* __restrict__ st_vz,
                                                      int* __restrict__ st_bin, R* __restrict__ acc) {
  const GlobalGrid<R> grid{acc};
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
    const long long i = i0 + j;
    const int s = species_of(i, p);
    const R q = p.sp_q[s];
    const R X0 = x0[3 * j];
    R v[3] = {v0[3 * j], v0[3 * j + 1], v0[3 * j + 2]};
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
      deposit_jx_startup(grid, xm, cm, cp, qj / p.dt, p);
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
  for (int b = lo; b < hi; ++b) mine += ((long long)bd.cur[0][b] + kBlk - 1) & ~(long long)(kBlk - 1);
  long long run = block_exclusive_scan<long long>(mine, &tot, sh_ll);
  for (int b = lo; b < hi; ++b) { bd.off[0][b] = run; run += ((long long)bd.cur[0][b] + kBlk - 1) & ~(long long)(kBlk - 1); }
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
      const R* q = slot_ptr(bd.rec[src], o + i);
      if (x_out) { x_out[3 * k] = node_pos(c, p) + q[0] * p.dx; x_out[3 * k + 1] = R(0); x_out[3 * k + 2] = R(0); }
      if (v_out) { v_out[3 * k] = q[kBlk]; v_out[3 * k + 1] = q[2 * kBlk]; v_out[3 * k + 2] = q[3 * kBlk]; }
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
      const R* q = slot_ptr(bd.rec[src], o + i);
      const double a = q[kBlk], b_ = q[2 * kBlk], c_ = q[3 * kBlk];
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
  size_t plan_smem_max = 0;

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
    bd.cap_total = (long long)((double)N * (1.0 + 3.0 * bd.slack)) + 96ll * bd.nb + 1024;
    bd.cap_total = (bd.cap_total + kBlk - 1) & ~(long long)(kBlk - 1);
    if (bd.cap_total >= (1ll << 40)) return e.fail(JIC_ERR_UNSUPPORTED, "too many particles for one GPU");
    bd.ov_cap = (int)std::min<long long>(std::max<long long>(N / 16, 1 << 16), 1ll << 28);
    bd.item_cap = (int)std::min<long long>(N / kSlowChunk + bd.nb + 16, 1ll << 30);
    bd.n_workers = n_sm * push_min_blocks<R>() * kPushWarps;
    {
      const bool periodic = dp.pbl == JIC_BC_PERIODIC && dp.pbr == JIC_BC_PERIODIC;
      // periodic: every bin takes the closed form (field_solver runs fix the reference's left-half-cell quirk up in place);
      // walls: the closed form's stencil (3 nodes for the moments, 3 faces more with field_solver) must stay on the grid
      bd.edge = periodic ? 0 : (dp.stag ? 3 : 2);
      if (dp.G < (dp.stag ? 10 : 8)) bd.edge = dp.G;  // tiny grids: the closed form's stencil would wrap onto itself
    }
    int rc;
    for (int k = 0; k < 2; ++k) {
      if ((rc = alloc(e, &bd.rec[k], 4 * (size_t)bd.cap_total))) return rc;
      if ((rc = alloc(e, &bd.off[k], bd.nb + 1)) || (rc = alloc(e, &bd.cnt[k], bd.nb)) || (rc = alloc(e, &bd.cur[k], bd.nb))) return rc;
      if ((rc = alloc(e, &bd.ov_bin[k], bd.ov_cap)) || (rc = alloc(e, &bd.ov_d[k], bd.ov_cap)) || (rc = alloc(e, &bd.ov_vx[k], bd.ov_cap)) ||
          (rc = alloc(e, &bd.ov_vy[k], bd.ov_cap)) || (rc = alloc(e, &bd.ov_vz[k], bd.ov_cap)))
        return rc;
    }
    if ((rc = alloc(e, &bd.item_bin, bd.item_cap)) || (rc = alloc(e, &bd.item_first, bd.item_cap)) || (rc = alloc(e, &bd.hdr, 1))) return rc;
    if ((rc = alloc(e, &dense, bd.nb + 1)) || (rc = alloc(e, &bd.psync, 1))) return rc;
    {
      const char* env = getenv("JIC_PLAN_MC");  // "0" keeps the single-CTA plan
      plan_ctas = (env && env[0] == '0') ? 0 : std::min(kPlanMaxCtas, (bd.nb + kPlanMcThreads - 1) / kPlanMcThreads);
      if (plan_ctas < 2) plan_ctas = 0;
    }
    (void)prm;
    {
      // Shared-memory carve-out: three CTAs of the fp64 kernel (43.3 KB static + 1 KB reserved each) fit the 132 KB configuration,
      // which is what the driver picks by itself; the remaining 124 KB of L1 matter (measured: 6 % slower with the 164 KB
      // configuration, 20 % with the maximum).  JIC_PUSH_CARVEOUT=<per cent of 228 KB> overrides for experiments.
      if (const char* env = getenv("JIC_PUSH_CARVEOUT")) {
        const int carve = atoi(env);
        cudaFuncSetAttribute(k_push<R, false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(k_push<R, true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(k_push<R, false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(k_push<R, true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
      }
    }
    {
      int dev = 0, max_smem = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
      plan_smem_max = max_smem > 4096 ? (size_t)max_smem - 4096 : 0;
      if ((size_t)2 * bd.nb * sizeof(int) > 48 * 1024 && (size_t)2 * bd.nb * sizeof(int) <= plan_smem_max)
        cudaFuncSetAttribute(k_plan<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)2 * bd.nb * sizeof(int)));
    }
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

  // initial binning: staging (in buffer 1, which is free until the first push) -> histogram -> exact layout -> scatter.
  // Three parts, so that a pipelined host upload can feed the first kernel chunk by chunk.
  int* st_bin = nullptr;
  int start_begin(Engine& e, const DevParams<R>& dp, cudaStream_t st) {
    cudaMemsetAsync(bd.hdr, 0, sizeof(PlanHeader), st);
    cudaMemsetAsync(bd.cur[0], 0, sizeof(unsigned) * bd.nb, st);
    st_bin = nullptr;
    cudaError_t ce = cudaMallocAsync((void**)&st_bin, sizeof(int) * (size_t)(dp.N ? dp.N : 1), st);
    if (ce != cudaSuccess) return e.fail(JIC_ERR_CUDA, format("staging allocation: %s", cudaGetErrorString(ce)));
    return JIC_OK;
  }
  int start_chunk(Engine& e, const DevParams<R>& dp, const R* x0, const R* v0, long long i0, long long n, R* acc, cudaStream_t st) {
    // buffer 1 is free until the first push: use it as four linear staging arrays
    R* sx = bd.rec[1]; R* svx = sx + bd.cap_total; R* svy = svx + bd.cap_total; R* svz = svy + bd.cap_total;
    k_start_binned<R><<<grid_for(n, 256, 8), 256, 0, st>>>(dp, bd, x0, v0, i0, n, sx, svx, svy, svz, st_bin, acc);
    e.launches += 1;
    return JIC_OK;
  }
  int start_end(Engine& e, const DevParams<R>& dp, cudaStream_t st) {
    R* sx = bd.rec[1]; R* svx = sx + bd.cap_total; R* svy = svx + bd.cap_total; R* svz = svy + bd.cap_total;
    k_first_layout<R><<<1, 1024, 0, st>>>(bd);
    k_scatter_binned<R><<<grid_for(dp.N, 256, 8), 256, 0, st>>>(dp, bd, sx, svx, svy, svz, st_bin);
    cudaFreeAsync(st_bin, st);
    st_bin = nullptr;
    e.launches += 2;
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) return e.fail(JIC_ERR_CUDA, format("binned start: %s", cudaGetErrorString(ce)));
    first_plan = true;
    return JIC_OK;
  }
  int start(Engine& e, const DevParams<R>& dp, const R* x0, const R* v0, R* acc, cudaStream_t st) {
    int rc = start_begin(e, dp, st);
    if (rc == JIC_OK) rc = start_chunk(e, dp, x0, v0, 0, dp.N, acc, st);
    return rc ? rc : start_end(e, dp, st);
  }
  bool first_plan = true;

  // runs after every push (and after the start-up scatter), concurrently with the field kernel: plan the next push
  int plan_ctas = 0;  // > 0: k_plan_mc on that many CTAs
  int plan(Engine& e, const DevParams<R>& dp, cudaStream_t st) {
    const size_t sm = (size_t)2 * bd.nb * sizeof(int);
    if (plan_ctas > 0) k_plan_mc<R><<<plan_ctas, kPlanMcThreads, 0, st>>>(bd, dp.G, first_plan ? 1 : 0);
    else if (sm <= plan_smem_max) k_plan<R, true><<<1, 1024, sm, st>>>(bd, dp.G, first_plan ? 1 : 0);
    else k_plan<R, false><<<1, 1024, 0, st>>>(bd, dp.G, first_plan ? 1 : 0);
    first_plan = false;
    e.launches += 1;
    return JIC_OK;
  }

  int step(Engine& e, const DevParams<R>& dp, const R* F, R* acc, cudaStream_t st) {
    const int g = n_sm * push_min_blocks<R>();
    if (dp.stag) {
      if (dp.relativistic) k_push<R, true, true><<<g, kPushThreads, 0, st>>>(dp, bd, F, acc);
      else k_push<R, false, true><<<g, kPushThreads, 0, st>>>(dp, bd, F, acc);
    } else {
      if (dp.relativistic) k_push<R, true, false><<<g, kPushThreads, 0, st>>>(dp, bd, F, acc);
      else k_push<R, false, false><<<g, kPushThreads, 0, st>>>(dp, bd, F, acc);
    }
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
