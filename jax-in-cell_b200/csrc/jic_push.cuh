// K1  the binned push (sm_100a): gather -> Boris -> move -> deposit -> re-bin in one pass over a (species, cell) bin.
// Included by jic_binned.cuh (store layout, slow_tail, warp_sum are defined there).
//
// Design, each point answering a counter of an earlier version's ncu profile (profiles/):
//
//  1. WARPS ARE THE WORKERS.  Every warp pulls work items (a run of <= chunk slots of one bin range) from the queue on its
//     own, owns a private shared-memory input ring, private mbarriers and private output staging, and flushes its own
//     deposit: no __syncthreads anywhere in the kernel, no producer/consumer imbalance inside a CTA.
//
//  2. PARTICLE LOADS ARE OFF THE INSTRUCTION STREAM.  The store keeps particles in 1 KiB blocks of 32
//     ([d x32][vx x32][vy x32][vz x32]); lane 0 streams whole blocks into the ring with 1-D bulk async copies
//     (cp.async.bulk.shared.global completing on an mbarrier: SASS UBLKCP + SYNCS).  Bytes in flight per SM =
//     (stages - 1) x stage bytes x resident warps, independent of registers; a particle costs four conflict-free LDS.64 at
//     immediate offsets instead of four LDG with 64-bit address arithmetic.
//
//  3. DEPOSITION USES MOMENTS, NOT PER-NODE WEIGHTS.  For offsets t in (-3/2, 3/2) from the bin's node c, every S2 weight
//     on nodes c-2..c+2 -- and every cumulative weight the charge-conserving J_x needs -- is a linear combination of
//     1, t, t^2, P(t) = max(t - 1/2, 0)^2, N(t) = max(-t - 1/2, 0)^2  (truncated-power form of the quadratic B-spline; the
//     knots at +-1/2 are the only ones inside the interval).  A lane accumulates 18 such sums (weighted by 1, v_y, v_z for
//     rho, J_y, J_z at x_{n+1}; by the displacement for J_x) with no selects and no branches; the 19 node values are formed
//     once per item from the warp totals.  With y = t - 1/2, 4 P = (y + |y|)^2: two additions (|.| is an operand modifier).
//         w(c-2) = N/2                          C(c-2) = N/2                      (C = cumulative weight up to the node,
//         w(c-1) = ((t-1/2)^2 - P - 3N)/2       C(c-1) = ((t-1/2)^2 - P)/2 - N     J_x(node) = -(q/dt) sum [C(t_new) - C(t_old)],
//         w(c)   = 3/4 - t^2 + 3(N+P)/2         C(c)   = (3/2-t)^2/2 - (t-1/2)^2 + P + N/2        P(t_old) = N(t_old) = 0)
//         w(c+1) = ((t+1/2)^2 - N - 3P)/2       C(c+1) = 1 - P/2
//         w(c+2) = P/2
//
//  4. RE-BINNING: STAYERS KEEP THEIR LANE, MOVERS ARE DEFERRED (round 2).  Measured with the re-binning compiled out, the
//     kernel runs at the copy peak (0.97 ms per 1e8 particles, profiles/r02_push_no_rebinning_floor.txt); the round-1
//     re-binning (a cursor atomic per block and destination, a deferred-store stash, scattered stores) cost 85 of its 229
//     instructions per particle and 0.34 ms.  Now:
//       * A particle that stays in its bin (|t_new| < 1/2: nine in ten) never moves between lanes.  A processed block of 32 goes
//         to the bin's BLOCK range straight from the registers, four 256-byte rows, as soon as its holes are filled.
//       * The holes that movers (and holes of the input) leave are filled from a POOL of already-processed stayers in shared
//         memory (at most 63): the lane with the r-th hole pops the r-th entry from the top.  When the pool holds fewer than 32
//         entries the next block becomes a donor: its stayers are pushed onto the pool (ballot ranks) instead of being written.
//         Every block written is therefore full; only the last one or two of a work item (the pool's remainder) carry holes
//         (d = NaN), which every reader skips at no cost -- all comparisons of the domain test |t| < 3/2 are false for NaN.
//       * Output blocks are claimed in runs of kPushRun blocks with one cursor atomic by lane 0, one run ahead of their use
//         (the round trip is > 1 us under load); fills and counts are warp-uniform, so every decision is a uniform branch.
//       * Movers go, with their unshifted offset, into a small ring in shared memory.  Every 32 of them are frozen: their slots
//         in the neighbours' SINGLE ranges are claimed (two atomics), and they are written when the next 32 are frozen -- ten
//         blocks later, when the claims have long arrived.
//     A bulk-copy flush (cp.async.bulk.global.shared::cta) of shared-memory staging blocks was measured first: the proxy fence
//     it needs compiles to MEMBAR.ALL.CTA, which waits for the pending claims, and staging every particle costs as many
//     instructions as round 1 did.
//
// Arithmetic follows jaxincell/_algorithms.py:40-66,90-92 (see jic_device.cuh for the per-function citations).  Particles
// that leave the closed form's domain (|t_new| >= 3/2, wall cells of non-periodic runs, full destination ranges) take the
// exact general code (slow_tail), which re-inserts them one by one into the destination's SINGLE range.
#pragma once

#ifndef JIC_EXPERIMENT_NOD
#define JIC_EXPERIMENT_NOD 0
#endif

namespace jic {

static_assert((JIC_PUSH_STAGES & (JIC_PUSH_STAGES - 1)) == 0 && (JIC_PUSH_STAGES_F32 & (JIC_PUSH_STAGES_F32 - 1)) == 0, "stage count must be a power of two");

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// 1-D bulk copy global -> shared; dst, src and bytes are multiples of 16
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(bar)
               : "memory");
}
// +1 / -1 / 0 cells as a real: the three doubles differ in their high word only
__device__ __forceinline__ double cell_shift(bool right, bool left, double) {
  return __hiloint2double(right ? 0x3ff00000 : (left ? (int)0xbff00000 : 0), 0);
}
__device__ __forceinline__ float cell_shift(bool right, bool left, float) { return right ? 1.f : (left ? -1.f : 0.f); }

template <typename R>
struct __align__(16) DestSlot {
  R* base;        // first block of the destination range
  unsigned cap;   // its capacity in slots
  int bin;        // range index (2 * bin + half)
};

// Packed fp32 arithmetic (sm_100: FFMA2 / FADD2 / FMUL2 work on two floats in a 64-bit register pair and take ONE issue slot): the fp32
// push is bound by instruction issue, and a lane carries two independent particles through phases B-C (push_stage_blocks = 2), so
// their arithmetic pairs up.  Same IEEE roundings as the scalar instructions: results are bit-identical to the scalar path.
#ifndef JIC_PUSH_F32X2
#define JIC_PUSH_F32X2 1
#endif
struct F32x2 { unsigned long long u; };
__device__ __forceinline__ F32x2 pk2(float lo, float hi) { F32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.u) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpk2(F32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v.u)); }
__device__ __forceinline__ F32x2 fma2(F32x2 a, F32x2 b, F32x2 c) { F32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.u) : "l"(a.u), "l"(b.u), "l"(c.u)); return r; }
__device__ __forceinline__ F32x2 add2(F32x2 a, F32x2 b) { F32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.u) : "l"(a.u), "l"(b.u)); return r; }
__device__ __forceinline__ F32x2 mul2(F32x2 a, F32x2 b) { F32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.u) : "l"(a.u), "l"(b.u)); return r; }
template <typename R> struct PushWarpPacked {};                      // (empty base: the fp64 layout is untouched)
template <> struct PushWarpPacked<float> { F32x2 coefp[24]; };       // coef[] with every value duplicated into both halves

// one warp's shared memory
constexpr int kPoolCap = 64;   // processed stayers waiting to fill holes (< 32 triggers a donor block, which adds at most 32)
constexpr int kMixCap = 128;   // ring of movers: 32 frozen (claimed, not yet written) + up to 63 collecting
template <typename R>
struct __align__(128) PushWarpSmem : PushWarpPacked<R> {
  R ring[push_stages<R>()][push_stage_blocks<R>()][kBlkElems];  // input: work-item blocks on their way in
  R pool[4][kPoolCap];                               // d, v_x, v_y, v_z of pooled stayers (a stack)
  R mix[4][kMixCap];                                 // t_new (unshifted), v_x, v_y, v_z of movers (a ring)
  R coef[24];                                        // per component k: [e0, e1, a2lo, a2hi, b0, b1, b2, -]
  R tot[24];
  unsigned long long full[push_stages<R>()];
  DestSlot<R> dest[3];                               // 0: the bin's block range; 1 / 2: the left / right neighbour's single range
};
template <typename R>
constexpr size_t push_smem_bytes() { return sizeof(PushWarpSmem<R>) * push_warps<R>(); }

// STAG (field_solver != 0): additionally rho(x_n) on the FACES c-3..c+2 (jaxincell/_algorithms.py:69-72), x_n = x_{n+1/2} - dt/2 v_n
// at offset ts from node c.  The face weights are the same B-spline seen from half a cell away, so their knots inside
// (-3/2, 3/2) sit at -1, 0, 1:  with Nm = max(-ts-1,0)^2, Z = max(ts,0)^2, Pp = max(ts-1,0)^2
//     W(c-3) = Nm/2                               W(c)   = ((ts+1)^2 - Nm - 3Z + 3Pp)/2
//     W(c-2) = (ts^2 - 3Nm - Z)/2                 W(c+1) = (Z - 3Pp)/2
//     W(c-1) = (1 - 2ts - 2ts^2 + 3Nm + 3Z - Pp)/2     W(c+2) = Pp/2
// Five more sums per lane, six more node values per item.  Bins next to the domain ends take the general path: there the
// reference drops the weight of faces -1 and -2 instead of wrapping it (make_cloud_faces).
template <typename R, bool REL, bool STAG>
__global__ void JIC_PUSH_BOUNDS(R) k_push(const __grid_constant__ DevParams<R> p, const __grid_constant__ BinDev<R> bd,
                                                                       const R* __restrict__ F, R* __restrict__ acc, RunControl* __restrict__ ctl) {
  constexpr int NS = push_stages<R>(), KB = push_stage_blocks<R>(), RUN = kPushRun;
  constexpr bool PACK = JIC_PUSH_F32X2 && std::is_same<R, float>::value && KB == 2 && !REL && !STAG;
  if (threadIdx.x == 0) atomicMin(&ctl->push_t0, global_timer_ns());
  constexpr unsigned kBlockBytes = kBlkElems * sizeof(R);
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(128) unsigned char push_smem_raw[];

  PlanHeader* hdr = bd.hdr;
  const int src = hdr->flip, dst = src ^ 1;
  const R* __restrict__ srec = bd.rec[src];
  const int n_items = hdr->n_items, chunk = hdr->chunk, tail_chunk = hdr->tail_chunk, tail_from = hdr->tail_from;
  const int warp = threadIdx.x >> 5;
  int lane;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));  // volatile: kept in a register instead of re-reading SR_TID in the loop
  const unsigned lt_mask = (1u << lane) - 1u;
  const int G = p.G;
  const R cells_per_v = p.dt * p.inv_dx;  // displacement in cells per unit velocity
  const R hole = hole_value(R(0));

  PushWarpSmem<R>& sm = reinterpret_cast<PushWarpSmem<R>*>(push_smem_raw)[warp];
  const R* my_ring = &sm.ring[0][0][0];
  const unsigned ring_u32 = smem_u32(my_ring), bar_u32 = smem_u32(&sm.full[0]);
  const R* coef = sm.coef;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) mbar_init(bar_u32 + 8u * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  unsigned gt = 0;  // ring slots consumed by this warp so far: slot = gt % NS, phase parity = (gt / NS) & 1

  for (;;) {
    // ---- next work item from the queue (dynamic: items differ in size)
    int item = 0;
    if (lane == 0) item = atomicAdd(&hdr->work, 1);
    item = __shfl_sync(FULL, item, 0);
    if (item >= n_items) break;
    const int range = bd.item_bin[item];  // 2 * bin + half
    const int first = bd.item_first[item];
    const int b = range >> 1;
    const int s = b / G, c = b - s * G;
    if (c < bd.edge || c > G - 1 - bd.edge) continue;  // a wall bin: every particle takes the general path (k_push_general)
    constexpr bool fast_bin = true;
    // STAG on a periodic domain: the three bins from which x_n can land in the left half cell (see the correction in phase C)
    const bool quirk_bin = STAG && bd.edge == 0 && (c == G - 1 || c <= 1);
    const R quirk_shift = c == G - 1 ? R(-0.5) : R(c) + R(0.5);
    const int n = min(fast_bin ? (b >= tail_from ? tail_chunk : chunk) : kSlowChunk, bd.cnt[src][range] - first);  // slots: whole blocks
    const int nblk = n >> 5, ngroups = (nblk + KB - 1) / KB;
    const R* item_rec = srec + ((bd.off[src][range] + first) >> 5) * (long long)kBlkElems;

    // lane 0 starts streaming the item at once; the other lanes set up the item-uniform tables meanwhile
    const unsigned gt0 = gt;
    auto load_group = [&](int g) {
      const unsigned slot = (gt0 + (unsigned)g) & (NS - 1);
      const unsigned bytes = (unsigned)min(KB, nblk - g * KB) * kBlockBytes;
      mbar_expect_tx(bar_u32 + 8u * slot, bytes);
      bulk_g2s(ring_u32 + slot * (KB * kBlockBytes), item_rec + (size_t)g * (KB * kBlkElems), bytes, bar_u32 + 8u * slot);
    };
    if (lane == 0) {
      const int pre = min(NS, ngroups);
      for (int g = 0; g < pre; ++g) load_group(g);
    }

    // ---- item-uniform gather polynomials.  Rows c..c+3 of the padded table are f[c-2], f[c-1], f[c], f[c+1].
    //   E lives on faces: for d < 0 the stencil is faces (c-2, c-1, c), for d >= 0 faces (c-1, c, c+1):
    //     E(d) = 1/2 (f[c-1]+f[c]) + d (f[c]-f[c-1]) + d^2 a2,   a2 = 1/2 (f[c-2]+f[c]) - f[c-1]  (d<0),  1/2 (f[c-1]+f[c+1]) - f[c]  (d>=0)
    //   B lives on centres (c-1, c, c+1):
    //     B(d) = 1/8 (b[c-1]+b[c+1]) + 3/4 b[c] + d/2 (b[c+1]-b[c-1]) + d^2 (1/2 (b[c-1]+b[c+1]) - b[c])
    //   Non-relativistic: pre-scaled by (q/m) dt/2.
    __syncwarp();  // the previous item's readers of coef / dest / tot are done
    if (lane < 24) {
      const int k = lane >> 3, w_ = lane & 7;
      R val = R(0);
      if (w_ < 7) {
        const R hs = REL ? R(1) : p.sp_qm[s] * p.half_dt;
        const R* f = F + (size_t)c * kFieldRow + k;
        if (w_ < 4) {
          const R f0 = __ldg(f), f1 = __ldg(f + kFieldRow), f2 = __ldg(f + 2 * kFieldRow), f3 = __ldg(f + 3 * kFieldRow);
          val = w_ == 0 ? R(0.5) * (f1 + f2) : w_ == 1 ? (f2 - f1) : w_ == 2 ? (R(0.5) * (f0 + f2) - f1) : (R(0.5) * (f1 + f3) - f2);
        } else {
          const R b1 = __ldg(f + kFieldRow + 3), b2 = __ldg(f + 2 * kFieldRow + 3), b3 = __ldg(f + 3 * kFieldRow + 3);
          val = w_ == 4 ? (R(0.125) * (b1 + b3) + R(0.75) * b2) : w_ == 5 ? (R(0.5) * (b3 - b1)) : (R(0.5) * (b1 + b3) - b2);
        }
        val *= hs;
      }
      sm.coef[lane] = val;
      if constexpr (PACK) sm.coefp[lane] = pk2(val, val);
    } else if (lane < 27) {
      const int k = lane - 24;
      const int bk = k == 0 ? b : (k == 1 ? s * G + (c == 0 ? G - 1 : c - 1) : s * G + (c == G - 1 ? 0 : c + 1));
      const int rk = 2 * bk + (k ? 1 : 0);  // own bin: block range; neighbours: single ranges
      const long long o = bd.off[dst][rk];
      sm.dest[k].base = bd.rec[dst] + (o >> 5) * (long long)kBlkElems;
      sm.dest[k].cap = (unsigned)(bd.off[dst][rk + 1] - o);
      sm.dest[k].bin = rk;
    }
    __syncwarp();

    // ---- output bookkeeping (all warp-uniform; claims are made by lane 0 and stay pending in its register until broadcast)
    unsigned pool_n = 0;                 // stayers in the pool
    unsigned mix_head = 0, mix_tail = 0, mix_frozen = 0;  // movers: ring positions; entries [tail, tail + frozen) have claims
    unsigned out_i = 0, runs_claimed = 0;  // output blocks written; runs of RUN blocks claimed
    unsigned claim_run = 0, claim_l = 0, claim_r = 0;     // lane 0: next output run; frozen movers' slots in the left / right range
    bool run_ok = false;                 // the current run lies inside the block range
    R* out_ptr = nullptr;                // this lane's slot in the next output block
    unsigned* const cur0 = bd.cur[dst] + sm.dest[0].bin;
#if JIC_EXPERIMENT_NOD
    R* nod_out = bd.rec[dst] + ((bd.off[dst][2 * b] + ((range & 1) ? bd.cnt[src][2 * b] : 0) + first) >> 5) * (long long)kBlkElems;
    if (lane == 0) atomicAdd(cur0, (unsigned)n);
#else
    if (fast_bin) {
      if (lane == 0) claim_run = atomicAdd(cur0, (unsigned)(RUN * kBlk));
      runs_claimed = 1;
    }
#endif
    // one full block (holes = NaN in d) -> the next slot of the claimed run.  `later` = upper bound of the particles that can still
    // be written after this block (pool + unprocessed input): decides whether another run is claimed now, one run ahead.
    auto emit_block = [&](R o_d, R o_v0, R o_v1, R o_v2, unsigned later) {
      if ((out_i & (unsigned)(RUN - 1)) == 0u) {
        const unsigned cl = __shfl_sync(FULL, claim_run, 0);
        const DestSlot<R> ds = sm.dest[0];
        run_ok = cl + (unsigned)(RUN * kBlk) <= ds.cap;
        out_ptr = ds.base + (size_t)(cl >> 5) * kBlkElems + lane;
        if (later > (unsigned)((RUN - 1) * kBlk)) {
          if (lane == 0) claim_run = atomicAdd(cur0, (unsigned)(RUN * kBlk));
          runs_claimed += 1;
        }
      }
      if (run_ok) {
        out_ptr[0] = o_d; out_ptr[kBlk] = o_v0; out_ptr[2 * kBlk] = o_v1; out_ptr[3 * kBlk] = o_v2;
      } else if (o_d == o_d) {
        store_slot(bd, dst, sm.dest[0].bin, 0xffffffffu, o_d, o_v0, o_v1, o_v2);  // -> overflow list (rare)
      }
      out_ptr += kBlkElems;
      out_i += 1;
    };
    // movers: write the frozen entries to the slots claimed when they were frozen / freeze the next `cnt` entries
    auto write_frozen = [&]() {
      const unsigned e = (mix_tail + (unsigned)lane) & (unsigned)(kMixCap - 1);
      const R t = sm.mix[0][e], o_v0 = sm.mix[1][e], o_v1 = sm.mix[2][e], o_v2 = sm.mix[3][e];
      const bool act = (unsigned)lane < mix_frozen;
      const bool right = act && t > R(0), left = act && !(t > R(0));
      const unsigned mr = __ballot_sync(FULL, right), ml = __ballot_sync(FULL, left);
      const unsigned base_r = __shfl_sync(FULL, claim_r, 0), base_l = __shfl_sync(FULL, claim_l, 0);
      if (act) {
        const DestSlot<R> ds = sm.dest[right ? 2 : 1];
        const unsigned slot = (right ? base_r : base_l) + __popc((right ? mr : ml) & lt_mask);
        const R dn = t - cell_shift(right, left, R(0));
        if (slot < ds.cap) {
          R* q = ds.base + (size_t)(slot >> 5) * kBlkElems + (slot & 31u);
          q[0] = dn; q[kBlk] = o_v0; q[2 * kBlk] = o_v1; q[3 * kBlk] = o_v2;
        } else {
          store_slot(bd, dst, ds.bin, 0xffffffffu, dn, o_v0, o_v1, o_v2);
        }
      }
      mix_tail += mix_frozen;
      mix_frozen = 0;
    };
    auto freeze = [&](unsigned cnt) {
      const unsigned e = (mix_tail + (unsigned)lane) & (unsigned)(kMixCap - 1);
      const bool right = (unsigned)lane < cnt && sm.mix[0][e] > R(0);
      const unsigned nr = __popc(__ballot_sync(FULL, right)), nl = cnt - nr;
      if (lane == 0) {
        if (nr) claim_r = atomicAdd(bd.cur[dst] + sm.dest[2].bin, nr);
        if (nl) claim_l = atomicAdd(bd.cur[dst] + sm.dest[1].bin, nl);
      }
      mix_frozen = cnt;
    };

    // moment accumulators (see the header): rho/J_y/J_z at the mid offset, J_x from the old and new offsets
    R r1 = 0, r2 = 0, rP = 0, rN = 0;
    R y0 = 0, y1 = 0, y2 = 0, yP = 0, yN = 0;
    R z0 = 0, z1 = 0, z2 = 0, zP = 0, zN = 0;
    R a1 = 0, a2 = 0, aP = 0, aN = 0;
    R s1 = 0, s2 = 0, sN = 0, sZ = 0, sP = 0;  // STAG
    int n_fast = 0;  // warp-uniform: particles of this item on the closed form

    for (int g = 0; g < ngroups; ++g, ++gt) {
      const unsigned slot = gt & (NS - 1);
      mbar_wait(bar_u32 + 8u * slot, (gt / NS) & 1u);
      const R* stage_in = my_ring + slot * (KB * kBlkElems);
      // The KB particles of a lane go through the arithmetic TOGETHER (phases A-C are straight-line code over kb, so the
      // compiler interleaves the independent FP64 dependency chains and shares the coefficient loads); the warp-level
      // bookkeeping follows per block in phase D.  Blocks past the end of the item hold stale data and are masked out.
      R d[KB], v[KB][3], u[KB], tn[KB], tm[KB], vx_old[KB], ts[KB];
      bool fast[KB], all_fast[KB];
      // ---- A. take the particles
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        d[kb] = stage_in[kb * kBlkElems + lane];
        v[kb][0] = stage_in[kb * kBlkElems + kBlk + lane]; v[kb][1] = stage_in[kb * kBlkElems + 2 * kBlk + lane]; v[kb][2] = stage_in[kb * kBlkElems + 3 * kBlk + lane];
        if (KB > 1 && kb > 0 && g * KB + kb >= nblk) d[kb] = hole;  // (the last group of an item with an odd number of blocks)
      }
      __syncwarp();  // every lane has taken its particles: the ring slot can be refilled
      if (lane == 0 && g + NS < ngroups) load_group(g + NS);
      // ---- B. gather (quadratics in d), velocity update, move
      if constexpr (PACK) {
        // the two particles of the lane side by side in 64-bit registers; the same operations in the same order as the loop below
        const F32x2* cp = sm.coefp;
        const int hi0 = d[0] >= R(0) ? 3 : 2, hi1 = d[1] >= R(0) ? 3 : 2;
        const F32x2 D = pk2(d[0], d[1]), NEG1 = pk2(-1.f, -1.f);
        F32x2 E2[3], B2[3], nB2[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          E2[k] = fma2(fma2(pk2(coef[8 * k + hi0], coef[8 * k + hi1]), D, cp[8 * k + 1]), D, cp[8 * k]);
          B2[k] = fma2(fma2(cp[8 * k + 6], D, cp[8 * k + 5]), D, cp[8 * k + 4]);
          nB2[k] = mul2(B2[k], NEG1);
        }
        const F32x2 vm0 = add2(pk2(v[0][0], v[1][0]), E2[0]), vm1 = add2(pk2(v[0][1], v[1][1]), E2[1]), vm2 = add2(pk2(v[0][2], v[1][2]), E2[2]);
        const F32x2 R0 = fma2(vm1, B2[2], fma2(vm2, nB2[1], vm0)), R1 = fma2(vm2, B2[0], fma2(vm0, nB2[2], vm1)),
                    R2 = fma2(vm0, B2[1], fma2(vm1, nB2[0], vm2));
        const F32x2 Rt = fma2(R0, B2[0], fma2(R1, B2[1], mul2(R2, B2[2])));
        const F32x2 den = fma2(B2[0], B2[0], fma2(B2[1], B2[1], fma2(B2[2], B2[2], pk2(1.f, 1.f))));
        float den0, den1;
        unpk2(den, den0, den1);
        const F32x2 inv = pk2(rcp_fast(den0), rcp_fast(den1));
        const F32x2 V0 = fma2(fma2(R1, B2[2], fma2(R2, nB2[1], fma2(Rt, B2[0], R0))), inv, E2[0]);
        const F32x2 V1 = fma2(fma2(R2, B2[0], fma2(R0, nB2[2], fma2(Rt, B2[1], R1))), inv, E2[1]);
        const F32x2 V2 = fma2(fma2(R0, B2[1], fma2(R1, nB2[0], fma2(Rt, B2[2], R2))), inv, E2[2]);
        const F32x2 U = mul2(V0, pk2(cells_per_v, cells_per_v));
        const F32x2 TN = add2(D, U), TM = fma2(pk2(0.5f, 0.5f), U, D);
        unpk2(V0, v[0][0], v[1][0]); unpk2(V1, v[0][1], v[1][1]); unpk2(V2, v[0][2], v[1][2]);
        unpk2(U, u[0], u[1]); unpk2(TN, tn[0], tn[1]); unpk2(TM, tm[0], tm[1]);
        fast[0] = fast_bin && (fabs(tn[0]) < R(1.5));
        fast[1] = fast_bin && (fabs(tn[1]) < R(1.5));
      } else {
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        R E[3], B[3];
        const int hi = d[kb] >= R(0) ? 3 : 2;
        if (STAG) { vx_old[kb] = v[kb][0]; ts[kb] = fma(R(-0.5) * cells_per_v, v[kb][0], d[kb]); }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          E[k] = fma(fma(coef[8 * k + hi], d[kb], coef[8 * k + 1]), d[kb], coef[8 * k]);
          B[k] = fma(fma(coef[8 * k + 6], d[kb], coef[8 * k + 5]), d[kb], coef[8 * k + 4]);
        }
        if (REL) {
          boris_velocity_relativistic(v[kb], E, B, p.sp_q[s], p.sp_m[s], p.dt);
        } else {
          // E, B already carry the factor (q/m) dt/2:  v- = v + E ; t = B ; v+ = (R x t + (R.t) t + R)/(1 + t.t) ; v = v+ + E
          const R vm0 = v[kb][0] + E[0], vm1 = v[kb][1] + E[1], vm2 = v[kb][2] + E[2];
          const R R0 = fma(vm1, B[2], fma(-vm2, B[1], vm0)), R1 = fma(vm2, B[0], fma(-vm0, B[2], vm1)), R2 = fma(vm0, B[1], fma(-vm1, B[0], vm2));
          const R Rt = fma(R0, B[0], fma(R1, B[1], R2 * B[2]));
          const R inv = rcp_fast(fma(B[0], B[0], fma(B[1], B[1], fma(B[2], B[2], R(1)))));
          v[kb][0] = fma(fma(R1, B[2], fma(-R2, B[1], fma(Rt, B[0], R0))), inv, E[0]);
          v[kb][1] = fma(fma(R2, B[0], fma(-R0, B[2], fma(Rt, B[1], R1))), inv, E[1]);
          v[kb][2] = fma(fma(R0, B[1], fma(-R1, B[0], fma(Rt, B[2], R2))), inv, E[2]);
        }
        // offsets from node c in cells: new offset tn, mid offset tm.  A hole (d = NaN) fails the domain test by itself.
        u[kb] = v[kb][0] * cells_per_v;
        tn[kb] = d[kb] + u[kb];
        tm[kb] = fma(R(0.5), u[kb], d[kb]);
        fast[kb] = fast_bin && (fabs(tn[kb]) < R(1.5));
        if (STAG) fast[kb] = fast[kb] && (fabs(ts[kb]) < R(1.5));
      }
      }
      // ---- C. deposit moments
      auto moments = [&](int kb) {
        // truncated powers 4 P(t) = (y + |y|)^2 with y = t - 1/2, 4 N(t) likewise with y = -t - 1/2
        const R tn_ = tn[kb], tm_ = tm[kb], vy = v[kb][1], vz = v[kb][2];
        const R pn_ = (tn_ - R(0.5)) + fabs(tn_ - R(0.5)), nn_ = (-tn_ - R(0.5)) + fabs(-tn_ - R(0.5));
        const R pm_ = (tm_ - R(0.5)) + fabs(tm_ - R(0.5)), nm_ = (-tm_ - R(0.5)) + fabs(-tm_ - R(0.5));
        const R Pm = pm_ * pm_, Nm = nm_ * nm_, tm2 = tm_ * tm_;
        a1 += u[kb]; a2 = fma(u[kb], tn_ + d[kb], a2); aP = fma(pn_, pn_, aP); aN = fma(nn_, nn_, aN);
        r1 += tm_; r2 += tm2; rP += Pm; rN += Nm;
        y0 += vy; y1 = fma(vy, tm_, y1); y2 = fma(vy, tm2, y2); yP = fma(vy, Pm, yP); yN = fma(vy, Nm, yN);
        z0 += vz; z1 = fma(vz, tm_, z1); z2 = fma(vz, tm2, z2); zP = fma(vz, Pm, zP); zN = fma(vz, Nm, zN);
        if (STAG) {
          const R t_ = ts[kb];
          if (quirk_bin) {
            // The reference does not wrap the face weights of a particle whose x_n lies in the left half cell [-L/2, g_0): faces
            // -2 and -1 are simply off its grid (make_cloud_faces).  The sums below wrap them onto faces G-2, G-1: take them back.
            const R xi = t_ + quirk_shift;  // x_n in cells from the left wall
            if (xi >= R(0) && xi < R(0.5)) {
              const R a_ = p.sp_q[s] * p.inv_dx;
              atomicAdd(acc + (size_t)G * kAccRow + (G - 2), -a_ * R(0.5) * (R(0.5) - xi) * (R(0.5) - xi));
              atomicAdd(acc + (size_t)G * kAccRow + (G - 1), -a_ * (R(0.75) - xi * xi));
            }
          }
          const R n_ = (-t_ - R(1)) + fabs(-t_ - R(1)), z_ = t_ + fabs(t_), p_ = (t_ - R(1)) + fabs(t_ - R(1));
          s1 += t_; s2 = fma(t_, t_, s2); sN = fma(n_, n_, sN); sZ = fma(z_, z_, sZ); sP = fma(p_, p_, sP);
        }
      };
      bool group_fast = true;
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) { all_fast[kb] = __all_sync(FULL, fast[kb]); group_fast = group_fast && all_fast[kb]; }
      if (group_fast) {  // the common case, warp-uniform: one straight line over the KB particles of a lane
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) moments(kb);
      } else {
#pragma unroll
        for (int kb = 0; kb < KB; ++kb)
          if (fast[kb]) moments(kb);
      }
#if JIC_EXPERIMENT_NOD
      // TIMING EXPERIMENT ONLY (wrong physics): every particle keeps its bin and slot -- the cost of the kernel without re-binning
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        if (g * KB + kb < nblk) {
          R* q = nod_out + (size_t)(g * KB + kb) * kBlkElems + lane;
          q[0] = d[kb]; q[kBlk] = v[kb][0]; q[2 * kBlk] = v[kb][1]; q[3 * kBlk] = v[kb][2];
        }
        n_fast += __popc(__ballot_sync(FULL, fast[kb]));
      }
      continue;
#endif
      // ---- D. per block: stayers keep their lane; holes are filled from the pool; movers are parked in the ring
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        const bool fs = fast[kb];
        const bool stay = fs && fabs(tn[kb]) < R(0.5);  // (false for a hole: NaN)
        const unsigned ms = __ballot_sync(FULL, stay);
        R o_d = tn[kb], o_v0 = v[kb][0], o_v1 = v[kb][1], o_v2 = v[kb][2];
        if (ms != FULL) {
          const bool mover = fs && !stay;
          const unsigned mm = __ballot_sync(FULL, mover);
          n_fast += __popc(ms | mm);
          if (mm) {
            if (mover) {
              const unsigned e = (mix_head + __popc(mm & lt_mask)) & (unsigned)(kMixCap - 1);
              sm.mix[0][e] = o_d; sm.mix[1][e] = o_v0; sm.mix[2][e] = o_v1; sm.mix[3][e] = o_v2;
            }
            mix_head += __popc(mm);
          }
          if (!all_fast[kb]) {
            // a particle (not a hole) off the closed form: its velocity is final; boundary, deposit and re-insertion are k_push_general's
            const bool general = !fs && d[kb] == d[kb];
            const unsigned mg = __ballot_sync(FULL, general);
            if (mg) {
              int base = 0;
              if (lane == 0) base = atomicAdd(&hdr->gen_n, __popc(mg));
              base = __shfl_sync(FULL, base, 0);
              const int k = base + __popc(mg & lt_mask);
              if (general) {
                if (k < bd.gen_cap) {
                  bd.gen_bin[k] = b; bd.gen_d[k] = d[kb]; bd.gen_vxold[k] = STAG ? vx_old[kb] : R(0);
                  bd.gen_vx[k] = v[kb][0]; bd.gen_vy[k] = v[kb][1]; bd.gen_vz[k] = v[kb][2];
                } else {
                  atomicExch(&hdr->error, 1);
                }
              }
            }
          }
        } else {
          n_fast += kBlk;
        }
        if (fast_bin) {
          if (pool_n < (unsigned)kBlk) {
            // donor block: its stayers go to the pool
            __syncwarp();
            if (stay) {
              const unsigned e = pool_n + __popc(ms & lt_mask);
              sm.pool[0][e] = o_d; sm.pool[1][e] = o_v0; sm.pool[2][e] = o_v1; sm.pool[3][e] = o_v2;
            }
            pool_n += __popc(ms);
          } else {
            if (ms != FULL) {  // fill the holes from the top of the pool
              __syncwarp();
              if (!stay) {
                const unsigned e = pool_n - 1u - __popc(~ms & lt_mask);
                o_d = sm.pool[0][e]; o_v0 = sm.pool[1][e]; o_v1 = sm.pool[2][e]; o_v2 = sm.pool[3][e];
              }
              pool_n -= __popc(~ms);
            }
            const unsigned done = ((unsigned)(g * KB + kb) + 1u) * (unsigned)kBlk;
            emit_block(o_d, o_v0, o_v1, o_v2, pool_n + (done < (unsigned)n ? (unsigned)n - done : 0u));
          }
          if (mix_head - mix_tail - mix_frozen >= (unsigned)kBlk) {  // 32 more movers: write the frozen ones, freeze these
            __syncwarp();
            if (mix_frozen) write_frozen();
            freeze((unsigned)kBlk);
          }
        }
      }
    }

    // ---- close the item's output: the pool's remainder (last block padded with holes), the movers, unused claims
    if (fast_bin && !JIC_EXPERIMENT_NOD) {
      __syncwarp();
      for (unsigned i0 = 0; i0 < pool_n; i0 += (unsigned)kBlk) {
        const unsigned e = i0 + (unsigned)lane;
        const bool ok = e < pool_n;
        const unsigned ec = ok ? e : 0u;
        const R o_d = ok ? sm.pool[0][ec] : hole;
        emit_block(o_d, sm.pool[1][ec], sm.pool[2][ec], sm.pool[3][ec], pool_n - min(pool_n, i0 + (unsigned)kBlk));
      }
      if (mix_frozen) write_frozen();
      while (mix_head != mix_tail) {
        freeze(min((unsigned)kBlk, mix_head - mix_tail));
        write_frozen();
      }
      // claimed but never written: the rest of the current run, and a whole run if the bound that claimed it was not reached
      if (run_ok)
        for (unsigned k = out_i & (unsigned)(RUN - 1); k != 0u && k < (unsigned)RUN; ++k) { out_ptr[0] = hole; out_ptr += kBlkElems; }
      if (runs_claimed * (unsigned)RUN >= out_i + (unsigned)RUN) {
        const unsigned cl = __shfl_sync(FULL, claim_run, 0);
        const DestSlot<R> ds = sm.dest[0];
        if (cl + (unsigned)(RUN * kBlk) <= ds.cap)
          for (int k = 0; k < RUN; ++k) ds.base[(size_t)((cl >> 5) + k) * kBlkElems + lane] = hole;
      }
      __syncwarp();
    }

    // ---- flush: warp totals of the moments -> 19 node values -> atomics on the raw (L2-resident) grid
    if (n_fast > 0) {
      R vals[18] = {r1, r2, rP, rN, y0, y1, y2, yP, yN, z0, z1, z2, zP, zN, a1, a2, aP, aN};
#pragma unroll
      for (int j = 0; j < 18; ++j) {
        const R tsum = warp_sum(vals[j]);
        if (lane == j) sm.tot[j] = tsum;
      }
      if (STAG) {
        R sv[5] = {s1, s2, sN, sZ, sP};
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const R tsum = warp_sum(sv[j]);
          if (lane == j) sm.tot[18 + j] = tsum;
        }
      }
      __syncwarp();
      if (STAG && lane >= 19 && lane < 25) {  // rho(x_n) on the faces c-3..c+2 (the truncated-power sums carry a factor 4)
        const R* tot = sm.tot;
        const R X0 = (R)n_fast, X1 = tot[18], X2 = tot[19], Nm = R(0.25) * tot[20], Z = R(0.25) * tot[21], Pp = R(0.25) * tot[22];
        const int o = lane - 19;
        R val = o == 0 ? Nm
              : o == 1 ? X2 - R(3) * Nm - Z
              : o == 2 ? X0 - R(2) * (X1 + X2) + R(3) * (Nm + Z) - Pp
              : o == 3 ? X2 + R(2) * X1 + X0 - Nm - R(3) * (Z - Pp)
              : o == 4 ? Z - R(3) * Pp
                       : Pp;
        val *= R(0.5) * p.sp_q[s] * p.inv_dx;
        if (val != R(0)) atomicAdd(acc + (size_t)G * kAccRow + mod_pos(c - 3 + o, G), val);
      }
      if (lane < 19) {
        const R* tot = sm.tot;
        const R cnt = (R)n_fast;
        const int j = lane;
        R val;
        int node, comp;
        if (j < 15) {
          const int f = j / 5, o = j - 5 * f;  // f: 0 rho, 1 J_y, 2 J_z ; o: node c-2+o   (P, N sums carry a factor 4)
          const int m = f == 0 ? 0 : (f == 1 ? 5 : 10);  // index of X1
          const R X0 = f == 0 ? cnt : tot[m - 1];
          const R X1 = tot[m], X2 = tot[m + 1], XP = tot[m + 2], XN = tot[m + 3];
          val = o == 0 ? R(0.125) * XN
              : o == 1 ? R(0.5) * (X2 - X1) + R(0.125) * (X0 - XP) - R(0.375) * XN
              : o == 2 ? R(0.75) * X0 - X2 + R(0.375) * (XN + XP)
              : o == 3 ? R(0.5) * (X2 + X1) + R(0.125) * (X0 - XN) - R(0.375) * XP
                       : R(0.125) * XP;
          node = c - 2 + o; comp = f == 0 ? 3 : f; val *= p.sp_q[s] * p.inv_dx;
        } else {
          const int o = j - 15;  // J_x on nodes c-2..c+1
          const R A1 = tot[14], A2 = tot[15], AP = tot[16], AN = tot[17];
          val = o == 0 ? R(0.125) * AN
              : o == 1 ? R(0.5) * (A2 - A1) - R(0.125) * AP - R(0.25) * AN
              : o == 2 ? R(-0.5) * (A2 + A1) + R(0.25) * AP + R(0.125) * AN
                       : R(-0.125) * AP;
          node = c - 2 + o; comp = 0; val *= -(p.sp_q[s] / p.dt);
        }
        if (val != R(0)) atomicAdd(acc + mod_pos(node, G) * kAccRow + comp, val);
      }
    }
  }

  __syncthreads();
  if (threadIdx.x == 0) atomicMax(&ctl->push_t1, global_timer_ns());
}

// ---------------------------------------------------------------------------------------------------------
// K1g  everything off the closed form, one thread per particle with the exact general code (slow_tail): global atomics on the raw
//      grid, re-insertion into the destination's single range.  Kept out of k_push so that its hot loop has no calls (the call's
//      register needs came on top of the 18 live accumulators: 203 registers wanted, 168 available, spills in the loop).
//        (a) the work items of wall bins (non-periodic runs; every bin on tiny grids), straight from the store;
//        (b) the particles k_push found off the closed form in this step (velocity already final);
//        (c) the particles that did not fit their range last step (overflow list of the source buffer).
//      On a periodic run at CFL <= 1 all three are empty and the kernel returns at once.
// ---------------------------------------------------------------------------------------------------------
constexpr int kGeneralThreads = 256;
static_assert(kGeneralThreads >= kSlowChunk, "one thread per slot of a wall-bin item");

template <typename R, bool REL>
__global__ void __launch_bounds__(kGeneralThreads) k_push_general(const __grid_constant__ DevParams<R> p, const __grid_constant__ BinDev<R> bd,
                                                                  const R* __restrict__ F, R* __restrict__ acc, RunControl* __restrict__ ctl) {
  PlanHeader* hdr = bd.hdr;
  const int src = hdr->flip, dst = src ^ 1;
  const int G = p.G;
  if (blockIdx.x == 0 && threadIdx.x == 0) {  // fold the push kernel's device-side duration into the running sum, re-arm the stamps
    if (ctl->push_t1 > ctl->push_t0) { ctl->push_ns += ctl->push_t1 - ctl->push_t0; ctl->push_launches += 1; }
    ctl->push_t0 = ~0ull; ctl->push_t1 = 0ull;
  }
  // (a) wall-bin items: CTAs stride over the item list (skipped on periodic runs: no wall bins)
  if (bd.edge > 0) {
    const int n_items = hdr->n_items;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int range = bd.item_bin[item];
      const int b = range >> 1;
      const int s = b / G, c = b - s * G;
      if (!(c < bd.edge || c > G - 1 - bd.edge)) continue;
      const int first = bd.item_first[item];
      const int n = min(kSlowChunk, bd.cnt[src][range] - first);
      const int i = threadIdx.x;
      if (i < n) {
        const R* q = slot_ptr(bd.rec[src], bd.off[src][range] + first + i);
        const R d = q[0];
        if (d == d) {
          const R x_old = node_pos(c, p) + d * p.dx;
          R v[3] = {q[kBlk], q[2 * kBlk], q[3 * kBlk]};
          const R vx_old = v[0];
          R E[3], B[3];
          gather_fields(F, x_old, p, E, B);
          if (REL) boris_velocity_relativistic(v, E, B, p.sp_q[s], p.sp_m[s], p.dt);
          else boris_velocity(v, E, B, p.sp_qm[s], p.dt);
          slow_tail(p, bd, dst, acc, s, x_old, vx_old, v[0], v[1], v[2]);
        }
      }
    }
  }
  // (b) this step's general-path list
  const int n_gen = min(hdr->gen_n, bd.gen_cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_gen; i += gridDim.x * blockDim.x) {
    const int b = bd.gen_bin[i];
    const int s = b / G, c = b - s * G;
    slow_tail(p, bd, dst, acc, s, node_pos(c, p) + bd.gen_d[i] * p.dx, bd.gen_vxold[i], bd.gen_vx[i], bd.gen_vy[i], bd.gen_vz[i]);
  }
  // (c) particles that did not fit their range last step
  const int n_ov = min(hdr->ov_n[src], bd.ov_cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_ov; i += gridDim.x * blockDim.x) {
    const int b = bd.ov_bin[src][i] >> 1;
    const int s = b / G, c = b - s * G;
    const R x_old = node_pos(c, p) + bd.ov_d[src][i] * p.dx;
    R v[3] = {bd.ov_vx[src][i], bd.ov_vy[src][i], bd.ov_vz[src][i]};
    const R vx_old = v[0];
    R E[3], B[3];
    gather_fields(F, x_old, p, E, B);
    if (REL) boris_velocity_relativistic(v, E, B, p.sp_q[s], p.sp_m[s], p.dt);
    else boris_velocity(v, E, B, p.sp_qm[s], p.dt);
    slow_tail(p, bd, dst, acc, s, x_old, vx_old, v[0], v[1], v[2]);
  }
}

}  // namespace jic
