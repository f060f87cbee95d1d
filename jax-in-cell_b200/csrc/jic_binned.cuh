// BINNED particle store (sm_100a): particles live in bins keyed by (species, cell of x_{n+1/2}) and are re-binned by
// the push kernel itself every step, so that for a whole work item
//   * the gather stencil is the same 4 grid rows  -> E(d), B(d) are quadratics in the in-cell offset d with item-uniform
//     coefficients (no per-particle field loads),
//   * the deposition stencil is the same 5 nodes  -> J_x, J_y, J_z, rho accumulate as moment sums in REGISTERS and reach the
//     L2-resident raw grid as 19 atomics per item (instead of ~15 atomics per particle),
//   * species constants (q w, q/m) are uniform,
// and a particle is stored as (d, v_x, v_y, v_z): its cell is implicit, d = (x - g_c)/dx in [-1/2, 1/2].
// Each particle's state crosses HBM once per step (4 reals in, 4 reals out).
//
// Memory layout: slots are grouped in BLOCKS of 32 (one warp); a block is [d x32][v_x x32][v_y x32][v_z x32], i.e. structure
// of arrays inside 1 KiB (fp64) records.  A warp reads a bin as a contiguous stream of whole blocks (one 1-D bulk async copy
// per pipeline stage) and every lane's four values sit at immediate offsets 0/256/512/768 B from one address.
//
// Every logical bin b = species * G + cell owns TWO slot ranges per buffer, indexed 2 b + half:
//   half 0, the BLOCK range: the particles that STAY in the bin, written by the push kernel in whole 32-slot blocks (runs of
//     kPushRun blocks per cursor atomic).  The last block a work item writes is partial and a run claimed ahead may stay empty:
//     unused slots hold d = NaN ("hole").  Readers skip holes for free: every comparison of the fast path's domain test
//     |t| < 3/2 is false for NaN.  (The initial scatter fills this range slot by slot.)
//   half 1, the SINGLE range: one slot per claim, dense: the particles that ARRIVE from other bins (movers of the fast path in
//     batches of up to 32, the general path one by one).
//   The plan pads the last block of every range with holes, so ranges are read as whole blocks.
// cnt[] and the cursors count SLOTS (holes included).
//
// The arithmetic is that of jaxincell/_algorithms.py:40-66,90-92 (see jic_device.cuh for the per-function citations); the
// "fast path" is the closed form of the reference's 6-node windowed prefix sum for a particle that moves by at most
// one cell and stays clear of non-periodic walls.  Everything else (multi-cell jumps, wall cells, overflowed bins) takes
// the exact general code of the INDEXED engine, particle by particle -- the deposit is additive, so paths can be mixed.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <type_traits>
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
// Launch shape of the push kernel.  Measured on B200 (profiles/r02_push_variants.txt): the kernel is latency-bound, so resident
// warps are what counts; fp64: one CTA of 16 warps per SM at 128 registers (two particles per lane in flight, 20 bytes of spills),
// fp32: five CTAs of 4 warps at 96 registers.
#ifndef JIC_PUSH_THREADS
#define JIC_PUSH_THREADS 512
#endif
#ifndef JIC_PUSH_MINBLOCKS
#define JIC_PUSH_MINBLOCKS 1       // 0: no CTA count in the launch bounds (register budget from -maxrregcount, for experiments)
#endif
#ifndef JIC_PUSH_THREADS_F32
#define JIC_PUSH_THREADS_F32 128
#endif
#ifndef JIC_PUSH_MINBLOCKS_F32
#define JIC_PUSH_MINBLOCKS_F32 5
#endif
#ifndef JIC_PUSH_STAGES
#define JIC_PUSH_STAGES 2          // ring slots per warp (power of two)
#endif
#ifndef JIC_PUSH_STAGES_F32
#define JIC_PUSH_STAGES_F32 4
#endif
#ifndef JIC_PUSH_STAGE_BLOCKS
#define JIC_PUSH_STAGE_BLOCKS 2    // 32-particle blocks per ring slot = particles a lane has in flight
#endif
#ifndef JIC_PUSH_STAGE_BLOCKS_F32
#define JIC_PUSH_STAGE_BLOCKS_F32 2
#endif
template <typename R>
__host__ __device__ constexpr int push_threads() { return sizeof(R) == 8 ? JIC_PUSH_THREADS : JIC_PUSH_THREADS_F32; }
template <typename R>
__host__ __device__ constexpr int push_min_blocks() { return sizeof(R) == 8 ? (JIC_PUSH_MINBLOCKS > 0 ? JIC_PUSH_MINBLOCKS : 1) : JIC_PUSH_MINBLOCKS_F32; }
template <typename R>
__host__ __device__ constexpr int push_stages() { return sizeof(R) == 8 ? JIC_PUSH_STAGES : JIC_PUSH_STAGES_F32; }
template <typename R>
__host__ __device__ constexpr int push_warps() { return push_threads<R>() / 32; }
// JIC_PUSH_MINBLOCKS=0: one CTA per SM whose register budget comes from -maxrregcount instead of the launch bounds
#if JIC_PUSH_MINBLOCKS > 0
#define JIC_PUSH_BOUNDS(R) __launch_bounds__(push_threads<R>(), push_min_blocks<R>())
#else
#define JIC_PUSH_BOUNDS(R) __launch_bounds__(push_threads<R>())
#endif
template <typename R>
__host__ __device__ constexpr int push_stage_blocks() { return sizeof(R) == 8 ? JIC_PUSH_STAGE_BLOCKS : JIC_PUSH_STAGE_BLOCKS_F32; }
#ifndef JIC_PUSH_RUN
#define JIC_PUSH_RUN 4             // output blocks claimed with one cursor atomic (power of two)
#endif
constexpr int kPushRun = JIC_PUSH_RUN;
#ifndef JIC_ITEMS_PER_WARP
#define JIC_ITEMS_PER_WARP 3       // work items per warp of the push kernel the plan aims for (load balance vs per-item overhead and holes)
#endif
#ifndef JIC_TAIL_SPLIT
#define JIC_TAIL_SPLIT 1           // the last eighth of the work queue is cut into quarter-size items (shorter kernel tail)
#endif

// worst-case holes one work item leaves in its bin's block range: the rest of its last run, one run claimed ahead that stayed
// empty, and the pool's partial last block
constexpr int kHolesPerWriter = (2 * kPushRun + 1) * kBlk;
// constant part of a block range's capacity: two items more than population / chunk + one block
constexpr int kBlockRangeConst = 2 * kHolesPerWriter + 32;

struct PlanHeader {
  int flip;            // which buffer is the SOURCE of the next push
  int n_items;         // work items of the next push
  int ov_n[2];         // entries in the overflow list of each buffer
  int error;           // sticky: 1 = overflow list full, 2 = capacity exhausted
  int chunk;           // particles per work item of the next push
  int work;            // dynamic work queue head (reset by k_plan)
  int tail_chunk;      // ... of the bins b >= tail_from (the end of the queue: smaller items, shorter kernel tail)
  int tail_from;
  int work2;           // work queue head of k_push_general (items of the wall bins)
  int gen_n;           // entries in the general-path list of this step (reset by k_plan)
  int gen_last;        // ... of the step before (diagnostics)
  long long n_stored;  // slots in use (holes included) + overflow entries in the source buffer
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
  int nb;                 // logical bins: n_species * G; slot ranges are indexed 2 b + half (0 = blocks, 1 = singles)
  long long cap_total;    // slots per buffer
  int ov_cap;             // overflow list capacity
  float slack;            // head-room fraction per neighbour
  R* rec[2];              // blocked particle records of each buffer: 4 * cap_total reals
  long long* off[2];      // [2 nb + 1] first slot of each range
  int* cnt[2];            // [2 nb]     slots in use in each range (a multiple of 32; holes included)
  unsigned* cur[2];       // [2 nb]     write cursors (count every attempt, also the overflowed ones)
  unsigned* slow0;        // [nb]       start-up only: particles per bin that will probably leave it in step 0
  // particles of fast bins that left the closed form's domain in this step (|t_new| >= 3/2): pushed by k_push, finished by k_push_general
  int* gen_bin; R* gen_d; R* gen_vxold; R* gen_vx; R* gen_vy; R* gen_vz; int gen_cap;
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

// what the d component of an unused slot holds
__device__ __forceinline__ double hole_value(double) { return __longlong_as_double(0x7ff8000000000000ll); }
__device__ __forceinline__ float hole_value(float) { return __int_as_float(0x7fc00000); }

template <typename R>
__device__ __forceinline__ R node_pos(int c, const DevParams<R>& p) { return p.g0 + R(c) * p.dx; }

// component a of slot k lives at slot_ptr(rec, k)[kBlk * a]
template <typename R>
__device__ __forceinline__ R* slot_ptr(R* rec, long long k) { return rec + ((k >> 5) << 7) + (k & 31); }

// Put one particle into slot `slot` of range `b` (= 2 * bin + half) of the destination buffer, or into the buffer's overflow
// list when the range is full.
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
  const int b = 2 * (species * p.G + c) + 1;  // the bin's SINGLE range
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
// block-wide exclusive prefix sum (used by the plan and the layout kernels)
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

// ---------------------------------------------------------------------------------------------------------
// K3  the plan, on NC CTAs: close the buffer that was just written, lay out the NEXT destination buffer, build the work-item
//     list, flip.  Every thread owns `per` contiguous logical bins (one at 8192 bins) with both of their ranges; every CTA sums
//     all cursors itself (the totals and the slack factor need no exchange), and the two prefix sums (slot offsets, work items)
//     are two-level: block scan, per-CTA totals through global memory, ONE grid-wide barrier (all CTAs are co-resident: the GPU
//     is otherwise idle at this point).
//     Capacities of the next buffer, from the attempts (cursor values) of this one -- pop = both ranges of a bin, arr = its
//     single range (what arrived last step; about as many leave):
//       block range   pop - arr / 2 + the holes its own work items can leave (kHolesPerWriter each)     [only stayers land here]
//       single range  arr + slack * (arr of 3 bins + pop)          (the whole population next to a non-periodic wall, where every
//                     particle takes the general path)
//     Before step 0 the start-up kernel's count of the particles about to leave each bin stands in for `arr`.
//     What does not fit lands in the overflow list and is re-inserted one step later.
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
  const int written = first_call ? h->flip : (h->flip ^ 1);  // buffer the last kernel wrote = source of the next push
  const int next = written ^ 1;                               // destination of the next push
  const unsigned* __restrict__ g_att = bd.cur[written];
  const long long* __restrict__ g_off = bd.off[written];
  const bool all_slow = bd.edge >= G;
  auto is_slow = [&](int c) { return c < bd.edge || c > G - 1 - bd.edge; };
  auto pop = [&](int b) -> long long { return (long long)g_att[2 * b] + (long long)g_att[2 * b + 1]; };
  auto left_of = [&](int b) { return (b % G) == 0 ? b + G - 1 : b - 1; };
  auto right_of = [&](int b) { return (b % G) == G - 1 ? b - (G - 1) : b + 1; };
  auto leaving = [&](int b) -> long long { return first_call ? (long long)bd.slow0[b] : (long long)g_att[2 * b + 1]; };
  auto arriving = [&](int b) -> long long {
    if (!first_call) return (long long)g_att[2 * b + 1];
    const long long l = bd.slow0[left_of(b)], r = bd.slow0[right_of(b)];
    return l > r ? l : r;
  };
  // totals: every CTA sums over all bins (a few KB from L2)
  long long mine = 0, mine_leave = 0, mine_arr = 0;
  for (int b = t; b < nb; b += nt) { mine += pop(b); mine_leave += leaving(b); mine_arr += arriving(b); }
  block_exclusive_scan<long long>(mine, &tot_ll, sh_ll);
  const long long n_total = tot_ll;
  block_exclusive_scan<long long>(mine_leave, &tot_ll, sh_ll);
  const long long t_leave = tot_ll;
  block_exclusive_scan<long long>(mine_arr, &tot_ll, sh_ll);
  const long long t_arr = tot_ll;
  long long want = n_total / ((long long)JIC_ITEMS_PER_WARP * (bd.n_workers > 0 ? bd.n_workers : 1));
  want = want < kMinChunk ? kMinChunk : (want > kMaxChunk ? kMaxChunk : want);
  const int kChunk = (int)((want + kChunkAlign - 1) / kChunkAlign) * kChunkAlign;
  int kTailChunk = ((kChunk / 4 + kChunkAlign - 1) / kChunkAlign) * kChunkAlign;  // the queue's tail: quarter-size items
  if (kTailChunk < kMinChunk) kTailChunk = kMinChunk < kChunk ? kMinChunk : kChunk;
  const int tail_from = JIC_TAIL_SPLIT ? nb - nb / 8 : nb;
  double f = bd.slack;
  {
    double room, per_f;
    if (all_slow) { room = (double)bd.cap_total - (double)n_total - 96.0 * nb; per_f = 3.0 * (double)n_total; }
    else {
      room = (double)bd.cap_total - ((double)n_total - 0.5 * (double)t_leave + (double)kHolesPerWriter * ((double)n_total / kTailChunk + 2.0 * nb) +
                                     (double)t_arr + 160.0 * nb);
      per_f = 3.0 * (double)t_arr + (double)n_total;
    }
    const double fmax = per_f > 0 ? room / per_f : 0.0;
    if (f > fmax) f = fmax;
    if (f < 0) { f = 0; if (t == 0 && cta == 0 && room < 0) atomicExch(&h->error, 2); }
  }
  // own logical bins: [lo, hi), contiguous over the whole grid of threads
  const int T = NC * nt, per = (nb + T - 1) / T;
  const int lo = min((cta * nt + t) * per, nb), hi = min(lo + per, nb);
  const int s_lo = lo / G, c_lo = lo - s_lo * G;
  auto chunk_of = [&](int b, int c) -> int { return is_slow(c) ? kSlowChunk : (b >= tail_from ? kTailChunk : kChunk); };
  auto caps_of = [&](int b, int c, long long& cb, long long& cs) {
    const int bl = c == 0 ? b + G - 1 : b - 1, br = c == G - 1 ? b - (G - 1) : b + 1;
    const long long p0 = pop(b);
    // block range: written by the bin's own (fast) work items only
    cb = is_slow(c) ? 0 : p0 - leaving(b) / 2 + (long long)kHolesPerWriter * (p0 / chunk_of(b, c) + 2) + 32;
    const bool near_wall = bd.edge > 0 && (c <= bd.edge || c >= G - 1 - bd.edge);
    if (near_wall) {
      cs = p0 + (long long)(f * (double)(p0 + pop(bl) + pop(br))) + 32;
    } else {
      const long long a0 = arriving(b), a3 = a0 + arriving(bl) + arriving(br);
      cs = a0 + (long long)(f * (double)(a3 + p0)) + 32;
    }
    cb = (cb + kBlk - 1) & ~(long long)(kBlk - 1);  // ranges start on block boundaries
    cs = (cs + kBlk - 1) & ~(long long)(kBlk - 1);
  };
  // 1. close `written`: slots in use = min(cursor, capacity); everything beyond sits in its overflow list.  A range that was
  //    filled slot by slot is padded with holes up to a whole block.
  long long cap_sum = 0;
  int item_sum = 0;
  {
    int c = c_lo;
    for (int b = lo; b < hi; ++b) {
      const int ch = chunk_of(b, c);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int i = 2 * b + half;
        const long long o = g_off[i], cap = g_off[i + 1] - o, att = g_att[i];
        int cnt = (int)(att < cap ? att : cap);
        if (cnt & (kBlk - 1)) {
          const int full = (cnt + kBlk - 1) & ~(kBlk - 1);
          for (int k = cnt; k < full; ++k) slot_ptr(bd.rec[written], o + k)[0] = hole_value(R(0));
          cnt = full;
        }
        bd.cnt[written][i] = cnt;
        item_sum += (cnt + ch - 1) / ch;
      }
      long long cb, cs;
      caps_of(b, c, cb, cs);
      cap_sum += cb + cs;
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
  // 2. offsets and cursors of `next`, 3. work items over `written`
  {
    int c = c_lo;
    for (int b = lo; b < hi; ++b) {
      long long cb, cs;
      caps_of(b, c, cb, cs);
      bd.off[next][2 * b] = run;
      bd.off[next][2 * b + 1] = run + cb;
      run += cb + cs;
      bd.cur[next][2 * b] = 0u;
      bd.cur[next][2 * b + 1] = 0u;
      const int ch = chunk_of(b, c);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int n = bd.cnt[written][2 * b + half];
        for (int k = 0; k < n; k += ch) {
          if (it < bd.item_cap) { bd.item_bin[it] = 2 * b + half; bd.item_first[it] = k; }
          ++it;
        }
      }
      c = c + 1 == G ? 0 : c + 1;
    }
  }
  if (cta == 0 && t == 0) {
    bd.off[next][2 * nb] = all_caps;
    if (all_items > bd.item_cap || all_caps > bd.cap_total) atomicExch(&h->error, 2);
    h->n_items = all_items < bd.item_cap ? all_items : bd.item_cap;
    h->chunk = kChunk;
    h->tail_chunk = kTailChunk;
    h->tail_from = tail_from;
    h->work = 0;
    h->work2 = 0;
    h->gen_last = h->gen_n;
    h->gen_n = 0;
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
      atomicAdd(&bd.cur[0][2 * bin], 1u);  // histogram of the first layout: everything starts in the block ranges, slot by slot
      // how many will leave the bin in step 0 (the velocity barely changes in one step): sizes the single ranges of the first plan
      {
        const R t_new = (xp - node_pos(c, p)) * p.inv_dx + v[0] * p.dt * p.inv_dx;
        if (!(fabs(t_new) < R(0.5))) atomicAdd(&bd.slow0[bin], 1u);
      }
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
  const int t = threadIdx.x, nt = blockDim.x, nb = 2 * bd.nb;  // all ranges (the block ranges are empty)
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
    const unsigned slot = atomicAdd(&bd.cur[0][2 * b], 1u);
    store_slot(bd, 0, 2 * b, slot, (st_x[i] - node_pos(c, p)) * p.inv_dx, st_vx[i], st_vy[i], st_vz[i]);
  }
}

// The same scatter in two passes (large runs).  k_scatter_binned writes four 8-byte words per particle into four different 32-byte
// sectors of a random block: every one of them is a partial-sector write that the memory system turns into a read-modify-write
// (10.6 ms per 1e8 particles, 3x the time of the rest of the data movement of the start-up).  Here pass 1 writes ONE whole sector per
// particle -- the record (d, v_x, v_y, v_z) at its final slot number in a temporary array-of-records -- and pass 2 streams each range's
// records back (coalesced) and writes the blocks' rows (coalesced).  Slot order inside a bin differs from run to run either way (cursor
// atomics); the sums the push forms over a bin do not depend on it beyond rounding.
template <typename R>
struct alignas(4 * sizeof(R)) BinRecord { R d, vx, vy, vz; };

// The temporary lives in memory the store already owns: buffer 1 is the start-up's staging area (four linear arrays of cap_total reals of
// which the first N are used), so each array has a free tail.  The record array is cut into four pieces, one per tail (the last one
// shorter: the bin indices of the staging area sit at its end).  No allocation on the start-up path.
template <typename R>
struct RecordPieces {
  BinRecord<R>* piece[4];
  long long len012, len3;  // records per piece
  __host__ __device__ long long capacity() const { return 3 * len012 + len3; }
  __device__ __forceinline__ BinRecord<R>* at(long long k) const {
    const int p = (k >= len012) + (k >= 2 * len012) + (k >= 3 * len012);
    return piece[p] + (k - p * len012);
  }
};

template <typename R>
__global__ void __launch_bounds__(256) k_scatter_records(const DevParams<R> p, const BinDev<R> bd, const R* __restrict__ st_x,
                                                         const R* __restrict__ st_vx, const R* __restrict__ st_vy, const R* __restrict__ st_vz,
                                                         const int* __restrict__ st_bin, const RecordPieces<R> tmp) {
  const long long tmp_n = tmp.capacity();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.N; i += (long long)gridDim.x * blockDim.x) {
    const int b = st_bin[i];
    if (b < 0) continue;
    const int c = b % p.G;
    const unsigned slot = atomicAdd(&bd.cur[0][2 * b], 1u);
    const long long k = bd.off[0][2 * b] + slot;
    const R d = (st_x[i] - node_pos(c, p)) * p.inv_dx;
    if (k < bd.off[0][2 * b + 1] && k < tmp_n) *tmp.at(k) = BinRecord<R>{d, st_vx[i], st_vy[i], st_vz[i]};
    else atomicExch(&bd.hdr->error, 2);  // (cannot happen: k_first_layout sized the ranges from the same histogram)
  }
}

template <typename R>
__global__ void __launch_bounds__(256) k_records_to_blocks(const BinDev<R> bd, const RecordPieces<R> tmp) {
  // one warp per block of 32 slots, over all blocks of the first layout; slots past a range's population were preset to NaN (= holes)
  const int lane = threadIdx.x & 31;
  const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long n_blocks = bd.off[0][2 * bd.nb] >> 5;
  for (long long blk = w; blk < n_blocks; blk += nw) {
    const BinRecord<R> rec = *tmp.at(blk * kBlk + lane);
    R* q = bd.rec[0] + blk * (long long)kBlkElems + lane;
    q[0] = rec.d; q[kBlk] = rec.vx; q[2 * kBlk] = rec.vy; q[3 * kBlk] = rec.vz;
  }
}

// export / diagnostics over the current source buffer (holes, d = NaN, are skipped) ------------------------------
// live particles per range: one warp per range
template <typename R>
__global__ void __launch_bounds__(256) k_count_live(const BinDev<R> bd, int* live /* 2 nb */) {
  const int src = bd.hdr->flip;
  const int lane = threadIdx.x & 31, w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int i = w; i < 2 * bd.nb; i += nw) {
    const long long o = bd.off[src][i];
    int n = 0;
    for (int k = lane; k < bd.cnt[src][i]; k += 32) {
      const R d = slot_ptr(bd.rec[src], o + k)[0];
      n += d == d ? 1 : 0;
    }
    for (int s = 16; s; s >>= 1) n += __shfl_xor_sync(0xffffffffu, n, s);
    if (lane == 0) live[i] = n;
  }
}

template <typename R>
__global__ void __launch_bounds__(1024) k_dense_offsets(const BinDev<R> bd, const int* live, long long* dense /* 2 nb + 1 */) {
  __shared__ long long sh_ll[1024];
  __shared__ long long tot;
  const int t = threadIdx.x, nt = blockDim.x, nb = 2 * bd.nb;
  const int per = (nb + nt - 1) / nt, lo = min(t * per, nb), hi = min(lo + per, nb);
  long long mine = 0;
  for (int b = lo; b < hi; ++b) mine += live[b];
  long long run = block_exclusive_scan<long long>(mine, &tot, sh_ll);
  for (int b = lo; b < hi; ++b) { dense[b] = run; run += live[b]; }
  if (t == 0) dense[nb] = tot;
}

template <typename R>
__global__ void __launch_bounds__(256) k_export_binned(const DevParams<R> p, const BinDev<R> bd, const long long* dense, R* x_out, R* v_out, uint8_t* alive) {
  const int src = bd.hdr->flip;
  const long long n_bins = dense[2 * bd.nb];
  const int lane = threadIdx.x & 31, w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int i = w; i < 2 * bd.nb; i += nw) {  // one warp per range, block after block, live particles ranked by ballot
    const int c = (i >> 1) % p.G;
    const long long o = bd.off[src][i];
    long long k = dense[i];
    for (int j = lane; j < bd.cnt[src][i]; j += 32) {
      const R* q = slot_ptr(bd.rec[src], o + j);
      const R d = q[0];
      const bool ok = d == d;
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        const long long kk = k + __popc(m & ((1u << lane) - 1u));
        if (x_out) { x_out[3 * kk] = node_pos(c, p) + d * p.dx; x_out[3 * kk + 1] = R(0); x_out[3 * kk + 2] = R(0); }
        if (v_out) { v_out[3 * kk] = q[kBlk]; v_out[3 * kk + 1] = q[2 * kBlk]; v_out[3 * kk + 2] = q[3 * kBlk]; }
        if (alive) alive[kk] = 1;
      }
      k += __popc(m);
    }
  }
  const int n_ov = min(bd.hdr->ov_n[src], bd.ov_cap);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.N - n_bins; i += (long long)gridDim.x * blockDim.x) {
    const long long k = n_bins + i;
    if (i < n_ov) {
      const int c = (bd.ov_bin[src][i] >> 1) % p.G;
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
  for (int i = blockIdx.x; i < 2 * bd.nb; i += gridDim.x) {
    const double m = (double)p.sp_m[(i >> 1) / p.G];
    const long long o = bd.off[src][i];
    for (int j = threadIdx.x; j < bd.cnt[src][i]; j += blockDim.x) {
      const R* q = slot_ptr(bd.rec[src], o + j);
      if (q[0] == q[0]) {
        const double a = q[kBlk], b_ = q[2 * kBlk], c_ = q[3 * kBlk];
        acc += 0.5 * m * (a * a + b_ * b_ + c_ * c_);
      }
    }
  }
  const int n_ov = min(bd.hdr->ov_n[src], bd.ov_cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_ov; i += gridDim.x * blockDim.x) {
    const double m = (double)p.sp_m[(bd.ov_bin[src][i] >> 1) / p.G];
    const double a = bd.ov_vx[src][i], b_ = bd.ov_vy[src][i], c_ = bd.ov_vz[src][i];
    acc += 0.5 * m * (a * a + b_ * b_ + c_ * c_);
  }
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

// per-species kinetic energy of the buffer the push just WROTE into row hist_row of jic_outputs.kinetic_energy (one warp per range;
// enqueued after the push kernels and before the plan flips the buffers)
template <typename R>
__global__ void __launch_bounds__(256) k_kinetic_hist_binned(const DevParams<R> p, const BinDev<R> bd, const RunControl* ctl) {
  double* out = (double*)ctl->hist[6];
  if (!out) return;
  out += ctl->hist_row * p.n_species;
  const int dst = bd.hdr->flip ^ 1;
  const int lane = threadIdx.x & 31, w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int i = w; i < 2 * bd.nb; i += nw) {
    const long long o = bd.off[dst][i], cap = bd.off[dst][i + 1] - o;
    const long long used = min((long long)bd.cur[dst][i], cap);  // (the plan has not closed this buffer yet: cursors, not counts)
    const int s = (i >> 1) / p.G;
    double acc = 0.0;
    for (long long k = lane; k < used; k += 32) {
      const R* q = slot_ptr(bd.rec[dst], o + k);
      if (q[0] == q[0]) {
        const double a = q[kBlk], b_ = q[2 * kBlk], c_ = q[3 * kBlk];
        acc += a * a + b_ * b_ + c_ * c_;
      }
    }
    for (int sh = 16; sh; sh >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sh);
    if (lane == 0 && acc != 0.0) atomicAdd(out + s, 0.5 * (double)p.sp_m[s] * acc);
  }
  const int n_ov = min(bd.hdr->ov_n[dst], bd.ov_cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_ov; i += gridDim.x * blockDim.x) {
    const int s = (bd.ov_bin[dst][i] >> 1) / p.G;
    const double a = bd.ov_vx[dst][i], b_ = bd.ov_vy[dst][i], c_ = bd.ov_vz[dst][i];
    atomicAdd(out + s, 0.5 * (double)p.sp_m[s] * (a * a + b_ * b_ + c_ * c_));
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side of the store
// ---------------------------------------------------------------------------------------------------------
template <typename R>
struct BinnedStore {
  BinDev<R> bd;
  std::vector<void*> owned;
  long long* dense = nullptr;
  int* live = nullptr;
  int n_sm = 148;
  bool built = false;

  template <typename T>
  int alloc(Engine& e, T** ptr, size_t n, bool zero = true) {
    void* p = nullptr;
    cudaError_t ce = DevicePool::get().malloc(&p, (n ? n : 1) * sizeof(T));
    if (ce != cudaSuccess) return e.fail(JIC_ERR_CUDA, format("cudaMalloc(%zu bytes) for the particle bins: %s", n * sizeof(T), cudaGetErrorString(ce)));
    if (zero) cudaMemset(p, 0, (n ? n : 1) * sizeof(T));
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
    // both ranges of every bin (see k_plan_mc).  Typical need: 1.35 N (a tenth of the particles change bins per step); the worst
    // case, every particle on the general path every step, is 2 N; the holes come on top (kHolesPerWriter per work item).
    bd.cap_total = (long long)((double)N * (2.25 + (double)kHolesPerWriter / kMinChunk)) + (long long)(5 * kHolesPerWriter + 256) * bd.nb + 4096;
    bd.cap_total = (bd.cap_total + kBlk - 1) & ~(long long)(kBlk - 1);
    if (bd.cap_total >= (1ll << 40)) return e.fail(JIC_ERR_UNSUPPORTED, "too many particles for one GPU");
    bd.ov_cap = (int)std::min<long long>(std::max<long long>(N / 16, 1 << 16), 1ll << 28);
    bd.item_cap = (int)std::min<long long>(bd.cap_total / kSlowChunk + 2 * bd.nb + 16, 1ll << 30);
    bd.n_workers = n_sm * push_min_blocks<R>() * push_warps<R>();
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
      if ((rc = alloc(e, &bd.off[k], 2 * bd.nb + 1)) || (rc = alloc(e, &bd.cnt[k], 2 * bd.nb)) || (rc = alloc(e, &bd.cur[k], 2 * bd.nb))) return rc;
      if ((rc = alloc(e, &bd.ov_bin[k], bd.ov_cap)) || (rc = alloc(e, &bd.ov_d[k], bd.ov_cap)) || (rc = alloc(e, &bd.ov_vx[k], bd.ov_cap)) ||
          (rc = alloc(e, &bd.ov_vy[k], bd.ov_cap)) || (rc = alloc(e, &bd.ov_vz[k], bd.ov_cap)))
        return rc;
    }
    if ((rc = alloc(e, &bd.item_bin, bd.item_cap)) || (rc = alloc(e, &bd.item_first, bd.item_cap)) || (rc = alloc(e, &bd.hdr, 1))) return rc;
    bd.gen_cap = (int)std::min<long long>(N + 1024, 1ll << 30);  // (a run far above CFL 1 sends every particle through this list)
    if ((rc = alloc(e, &bd.gen_bin, bd.gen_cap)) || (rc = alloc(e, &bd.gen_d, bd.gen_cap)) || (rc = alloc(e, &bd.gen_vxold, bd.gen_cap)) ||
        (rc = alloc(e, &bd.gen_vx, bd.gen_cap)) || (rc = alloc(e, &bd.gen_vy, bd.gen_cap)) || (rc = alloc(e, &bd.gen_vz, bd.gen_cap)))
      return rc;
    if ((rc = alloc(e, &dense, 2 * bd.nb + 1)) || (rc = alloc(e, &live, 2 * bd.nb)) || (rc = alloc(e, &bd.psync, 1)) || (rc = alloc(e, &bd.slow0, bd.nb))) return rc;
    plan_ctas = std::max(1, std::min(kPlanMaxCtas, (bd.nb + kPlanMcThreads - 1) / kPlanMcThreads));
    (void)prm;
    {
      // The push kernel's shared memory (input ring + output staging per warp) is dynamic: more than 48 KB per CTA needs the opt-in.
      // JIC_PUSH_CARVEOUT=<per cent of 228 KB> overrides the carve-out the driver derives from it, for experiments.
      const char* env = getenv("JIC_PUSH_CARVEOUT");
      const int carve = env ? atoi(env) : -1;
      const void* fns[4] = {(const void*)k_push<R, false, false>, (const void*)k_push<R, true, false>, (const void*)k_push<R, false, true>,
                            (const void*)k_push<R, true, true>};
      for (const void* fn : fns) {
        cudaError_t ce = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)push_smem_bytes<R>());
        if (ce != cudaSuccess) return e.fail(JIC_ERR_CUDA, format("push kernel: %zu bytes of shared memory per CTA: %s", push_smem_bytes<R>(), cudaGetErrorString(ce)));
        if (carve >= 0) cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
      }
    }
    built = true;
    return JIC_OK;
  }

  void destroy() {
    for (void* p : owned) DevicePool::get().free(p);  // (the engine's destructor has synchronised the device)
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
  bool st_bin_owned = false;      // st_bin came from cudaMallocAsync (the free tail of buffer 1 was too short)
  RecordPieces<R> st_records{};   // capacity() == 0: one-pass scatter
  int start_begin(Engine& e, const DevParams<R>& dp, cudaStream_t st) {
    cudaMemsetAsync(bd.hdr, 0, sizeof(PlanHeader), st);
    cudaMemsetAsync(bd.cur[0], 0, sizeof(unsigned) * 2 * bd.nb, st);
    cudaMemsetAsync(bd.slow0, 0, sizeof(unsigned) * bd.nb, st);
    // scratch of the start-up inside buffer 1: [array a: N staged reals | free tail], a = 0..3; the bin indices go to the end of the last
    // tail, the record pieces of the two-pass scatter to the (aligned) starts of the four tails
    const long long free_tail = bd.cap_total - dp.N;  // reals per array
    const long long bin_reals = ((long long)sizeof(int) * dp.N + sizeof(R) - 1) / (long long)sizeof(R) + 8;
    R* const stage = bd.rec[1];
    st_bin = nullptr;
    st_bin_owned = false;
    st_records = RecordPieces<R>{};
    if (free_tail >= bin_reals + 8) {
      uintptr_t a = (uintptr_t)(stage + 4 * bd.cap_total - bin_reals);
      st_bin = (int*)((a + 15) & ~(uintptr_t)15);
      const char* env = getenv("JIC_SCATTER_RECORDS");  // 0: never, 1 (default): from 2^20 particles up, 2: whenever the tails hold it
      const int mode = env ? atoi(env) : 1;
      const long long len012 = (free_tail - 8) / 4, len3 = (free_tail - bin_reals - 16) / 4;
      const long long need = dp.N + 32ll * bd.nb + 32;  // ranges are padded to whole blocks: at most 31 more slots per block range
      if ((mode == 2 || (mode == 1 && dp.N >= (1ll << 20))) && len3 > 0 && 3 * len012 + len3 >= need) {
        for (int a4 = 0; a4 < 4; ++a4) {
          const uintptr_t b = (uintptr_t)(stage + a4 * bd.cap_total + dp.N), al = 4 * sizeof(R);
          st_records.piece[a4] = (BinRecord<R>*)((b + al - 1) / al * al);
        }
        st_records.len012 = len012; st_records.len3 = len3;
      }
    } else {
      cudaError_t ce = cudaMallocAsync((void**)&st_bin, sizeof(int) * (size_t)(dp.N ? dp.N : 1), st);
      if (ce != cudaSuccess) return e.fail(JIC_ERR_CUDA, format("staging allocation: %s", cudaGetErrorString(ce)));
      st_bin_owned = true;
    }
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
    if (st_records.capacity() > 0) {
      // two-pass scatter through whole-sector records (see k_scatter_records); all-ones words are NaNs: unclaimed slots read back as holes
      for (int a4 = 0; a4 < 4; ++a4)
        cudaMemsetAsync(st_records.piece[a4], 0xFF, sizeof(BinRecord<R>) * (size_t)(a4 < 3 ? st_records.len012 : st_records.len3), st);
      k_scatter_records<R><<<grid_for(dp.N, 256, 8), 256, 0, st>>>(dp, bd, sx, svx, svy, svz, st_bin, st_records);
      k_records_to_blocks<R><<<grid_for(dp.N + 32ll * bd.nb + 32, 256, 8), 256, 0, st>>>(bd, st_records);
      e.launches += 1;
    } else {
      k_scatter_binned<R><<<grid_for(dp.N, 256, 8), 256, 0, st>>>(dp, bd, sx, svx, svy, svz, st_bin);
    }
    if (st_bin_owned) cudaFreeAsync(st_bin, st);
    st_bin = nullptr;
    st_bin_owned = false;
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
  int plan_ctas = 1;
  int plan(Engine& e, const DevParams<R>& dp, cudaStream_t st) {
    k_plan_mc<R><<<plan_ctas, kPlanMcThreads, 0, st>>>(bd, dp.G, first_plan ? 1 : 0);
    first_plan = false;
    e.launches += 1;
    return JIC_OK;
  }

  int step(Engine& e, const DevParams<R>& dp, const R* F, R* acc, RunControl* ctl, cudaStream_t st) {
    const int g = n_sm * push_min_blocks<R>();
    const size_t sm = push_smem_bytes<R>();
    if (dp.stag) {
      if (dp.relativistic) k_push<R, true, true><<<g, push_threads<R>(), sm, st>>>(dp, bd, F, acc, ctl);
      else k_push<R, false, true><<<g, push_threads<R>(), sm, st>>>(dp, bd, F, acc, ctl);
    } else {
      if (dp.relativistic) k_push<R, true, false><<<g, push_threads<R>(), sm, st>>>(dp, bd, F, acc, ctl);
      else k_push<R, false, false><<<g, push_threads<R>(), sm, st>>>(dp, bd, F, acc, ctl);
    }
    // everything off the closed form: wall-bin items, this step's general-path list, last step's overflow list (usually all empty)
    if (dp.relativistic) k_push_general<R, true><<<n_sm * 4, kGeneralThreads, 0, st>>>(dp, bd, F, acc, ctl);
    else k_push_general<R, false><<<n_sm * 4, kGeneralThreads, 0, st>>>(dp, bd, F, acc, ctl);
    e.launches += 2;
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
    k_count_live<R><<<n_sm * 4, 256, 0, st>>>(bd, live);
    k_dense_offsets<R><<<1, 1024, 0, st>>>(bd, live, dense);
    k_export_binned<R><<<n_sm * 4, 256, 0, st>>>(dp, bd, dense, x, v, alive);
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) return e.fail(JIC_ERR_CUDA, format("binned export: %s", cudaGetErrorString(ce)));
    return JIC_OK;
  }

  int kinetic_hist(Engine& e, const DevParams<R>& dp, const RunControl* ctl, cudaStream_t st) {
    k_kinetic_hist_binned<R><<<n_sm * 8, 256, 0, st>>>(dp, bd, ctl);
    e.launches += 1;
    return JIC_OK;
  }

  int kinetic(Engine& e, const DevParams<R>& dp, double* out, cudaStream_t st) {
    k_kinetic_binned<R><<<n_sm * 4, 256, 0, st>>>(dp, bd, out);
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) return e.fail(JIC_ERR_CUDA, format("binned kinetic: %s", cudaGetErrorString(ce)));
    return JIC_OK;
  }

  long long extra_launches_per_step() const { return 2; }  // k_push_general, k_plan
};

}  // namespace jic
