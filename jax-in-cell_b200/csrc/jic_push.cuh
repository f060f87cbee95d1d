// K1  the binned push (sm_100a): gather -> Boris -> move -> deposit -> re-bin in one pass over a (species, cell) bin.
// Included by jic_binned.cuh (store layout, slow_tail, warp_sum are defined there).
//
// Design, each point answering a counter of the first version's ncu profile (profiles/r01_push_binned_64x6_*: 328 thread
// instructions per particle, 18 % occupancy, warps stalled on their own global loads, a __syncthreads per chunk phase):
//
//  1. WARPS ARE THE WORKERS.  Every warp pulls work items (a run of <= chunk particles of one bin) from the queue on its
//     own, owns a private shared-memory ring and private mbarriers, and flushes its own deposit: no __syncthreads anywhere
//     in the kernel, no producer/consumer imbalance inside a CTA.
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
//  4. RE-BINNING costs one cursor atomic per warp, destination and block (lanes 0..2 claim for stay / left / right), and
//     the store of a particle is issued one iteration after its claim so that the atomic's round trip is hidden.
//
// Arithmetic follows jaxincell/_algorithms.py:40-66,90-92 (see jic_device.cuh for the per-function citations).  Particles
// that leave the closed form's domain (|t_new| >= 3/2, wall cells of non-periodic runs, full destination bins) take the exact
// general code (slow_tail).
#pragma once

namespace jic {

static_assert((kPushStages & (kPushStages - 1)) == 0, "stage count must be a power of two");

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

template <typename R>
struct __align__(16) DestSlot {
  R* base;        // first block of the destination bin
  unsigned cap;   // its capacity in slots
  int bin;
};

template <typename R>
__device__ __forceinline__ void store_slot_at(const BinDev<R>& bd, int dst, const DestSlot<R>& ds, unsigned slot, R d, R vx, R vy, R vz) {
  if (slot < ds.cap) {
    R* q = ds.base + (size_t)(slot >> 5) * kBlkElems + (slot & 31u);
    q[0] = d; q[kBlk] = vx; q[2 * kBlk] = vy; q[3 * kBlk] = vz;
  } else {
    store_slot(bd, dst, ds.bin, slot, d, vx, vy, vz);  // -> overflow list
  }
}

// STAG (field_solver != 0): additionally rho(x_n) on the FACES c-3..c+2 (jaxincell/_algorithms.py:69-72), x_n = x_{n+1/2} - dt/2 v_n
// at offset ts from node c.  The face weights are the same B-spline seen from half a cell away, so their knots inside
// (-3/2, 3/2) sit at -1, 0, 1:  with Nm = max(-ts-1,0)^2, Z = max(ts,0)^2, Pp = max(ts-1,0)^2
//     W(c-3) = Nm/2                               W(c)   = ((ts+1)^2 - Nm - 3Z + 3Pp)/2
//     W(c-2) = (ts^2 - 3Nm - Z)/2                 W(c+1) = (Z - 3Pp)/2
//     W(c-1) = (1 - 2ts - 2ts^2 + 3Nm + 3Z - Pp)/2     W(c+2) = Pp/2
// Five more sums per lane, six more node values per item.  Bins next to the domain ends take the general path: there the
// reference drops the weight of faces -1 and -2 instead of wrapping it (make_cloud_faces).
template <typename R, bool REL, bool STAG>
__global__ void __launch_bounds__(kPushThreads, push_min_blocks<R>()) k_push(const __grid_constant__ DevParams<R> p, const __grid_constant__ BinDev<R> bd,
                                                                       const R* __restrict__ F, R* __restrict__ acc) {
  constexpr int NW = kPushWarps, NS = kPushStages, KB = kPushStageBlocks;
  constexpr unsigned kBlockBytes = kBlkElems * sizeof(R);
  __shared__ __align__(128) R ring[NW][NS][KB][kBlkElems];
  __shared__ __align__(8) unsigned long long full[NW][NS];
  __shared__ __align__(16) R coef_s[NW][24];  // per component k: [e0, e1, a2lo, a2hi, b0, b1, b2, -]
  __shared__ DestSlot<R> dest_s[NW][3];
  __shared__ R tot_s[NW][24];
  __shared__ __align__(16) R stash[NW][KB][kBlkElems];  // re-binned particles waiting for their claimed slots

  PlanHeader* hdr = bd.hdr;
  const int src = hdr->flip, dst = src ^ 1;
  const R* __restrict__ srec = bd.rec[src];
  const int n_items = hdr->n_items, chunk = hdr->chunk;
  const int warp = threadIdx.x >> 5;
  int lane;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));  // volatile: kept in a register instead of re-reading SR_TID in the loop
  const unsigned lt_mask = (1u << lane) - 1u;
  const int G = p.G;
  const R cells_per_v = p.dt * p.inv_dx;  // displacement in cells per unit velocity

  const R* my_ring = &ring[warp][0][0][0];
  const unsigned ring_u32 = smem_u32(my_ring), bar_u32 = smem_u32(&full[warp][0]);
  const R* coef = coef_s[warp];
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
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= n_items) break;
    const int b = bd.item_bin[item];
    const int first = bd.item_first[item];
    const int s = b / G, c = b - s * G;
    const bool fast_bin = c >= bd.edge && c <= G - 1 - bd.edge;  // (k_plan makes the items of the other bins small)
    // STAG on a periodic domain: the three bins from which x_n can land in the left half cell (see the correction in phase C)
    const bool quirk_bin = STAG && bd.edge == 0 && (c == G - 1 || c <= 1);
    const R quirk_shift = c == G - 1 ? R(-0.5) : R(c) + R(0.5);
    const int n = min(fast_bin ? chunk : kSlowChunk, bd.cnt[src][b] - first);
    const int nblk = (n + kBlk - 1) / kBlk, ngroups = (nblk + KB - 1) / KB;
    const R* item_rec = srec + ((bd.off[src][b] + first) >> 5) * (long long)kBlkElems;

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
    __syncwarp();  // the previous item's readers of coef_s / dest_s / tot_s are done
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
      coef_s[warp][lane] = val;
    } else if (lane < 27) {
      const int k = lane - 24;
      const int bk = k == 0 ? b : (k == 1 ? s * G + (c == 0 ? G - 1 : c - 1) : s * G + (c == G - 1 ? 0 : c + 1));
      const long long o = bd.off[dst][bk];
      dest_s[warp][k].base = bd.rec[dst] + (o >> 5) * (long long)kBlkElems;
      dest_s[warp][k].cap = (unsigned)(bd.off[dst][bk + 1] - o);
      dest_s[warp][k].bin = bk;
    }
    __syncwarp();

    // lanes 0..2 claim slots for the destinations stay / left / right
    unsigned* my_cursor = bd.cur[dst] + dest_s[warp][lane < 3 ? lane : 0].bin;

    // moment accumulators (see the header): rho/J_y/J_z at the mid offset, J_x from the old and new offsets
    R r1 = 0, r2 = 0, rP = 0, rN = 0;
    R y0 = 0, y1 = 0, y2 = 0, yP = 0, yN = 0;
    R z0 = 0, z1 = 0, z2 = 0, zP = 0, zN = 0;
    R a1 = 0, a2 = 0, aP = 0, aN = 0;
    R s1 = 0, s2 = 0, sN = 0, sZ = 0, sP = 0;  // STAG
    int n_slow = 0;  // warp-uniform: particles of this item that took the general path

    // A particle is STORED one group (KB blocks) after its slot was claimed, so that the cursor atomic's round trip
    // (> 1 us under load: profiles/r01_push_warpworkers_*) overlaps a whole group's arithmetic instead of stalling the warp.
    // Until then its new state waits in the lane's own slot of a shared-memory stash (no cross-lane traffic, no conflicts)
    // and only (destination, rank) and the claimed base stay in registers.  q_meta < 0 = nothing pending.
    int q_meta[KB];
    unsigned q_claim[KB];
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) { q_meta[kb] = -1; q_claim[kb] = 0; }
    R* my_stash = &stash[warp][0][0] + lane;
    auto retire = [&](int kb) {
      const int kind = q_meta[kb] < 0 ? 0 : (q_meta[kb] & 3);
      const unsigned base_slot = __shfl_sync(0xffffffffu, q_claim[kb], kind);
      const DestSlot<R> ds = dest_s[warp][kind];
      const unsigned sl_ = base_slot + ((unsigned)q_meta[kb] >> 2);
      const bool pending = q_meta[kb] >= 0, fits = pending && sl_ < ds.cap;
      const R* st_ = my_stash + kb * kBlkElems;
      const R o_d = st_[0], o_v0 = st_[kBlk], o_v1 = st_[2 * kBlk], o_v2 = st_[3 * kBlk];
      if (fits) {
        R* q = ds.base + (size_t)(sl_ >> 5) * kBlkElems + (sl_ & 31u);
        q[0] = o_d; q[kBlk] = o_v0; q[2 * kBlk] = o_v1; q[3 * kBlk] = o_v2;
      }
      if (__any_sync(0xffffffffu, pending && !fits)) {
        if (pending && !fits) store_slot(bd, dst, ds.bin, sl_, o_d, o_v0, o_v1, o_v2);  // -> overflow list
      }
    };

    for (int g = 0; g < ngroups; ++g, ++gt) {
      const unsigned slot = gt & (NS - 1);
      mbar_wait(bar_u32 + 8u * slot, (gt / NS) & 1u);
      const R* stage = my_ring + slot * (KB * kBlkElems);
      // The KB particles of a lane go through the arithmetic TOGETHER (phases A-C are straight-line code over kb, so the
      // compiler interleaves the independent FP64 dependency chains and shares the coefficient loads); the warp-level
      // bookkeeping (votes, claims, stash) follows per block in phase D.  Blocks past the end of the item run with every lane
      // invalid: no early exit, the loop body stays one straight line.
      R d[KB], v[KB][3], u[KB], tn[KB], tm[KB], vx_old[KB], ts[KB];
      bool valid[KB], fast[KB], all_fast[KB];
      // ---- A. take the particles
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        valid[kb] = (g * KB + kb) * kBlk + lane < n;
        d[kb] = stage[kb * kBlkElems + lane];
        v[kb][0] = stage[kb * kBlkElems + kBlk + lane]; v[kb][1] = stage[kb * kBlkElems + 2 * kBlk + lane]; v[kb][2] = stage[kb * kBlkElems + 3 * kBlk + lane];
      }
      __syncwarp();  // every lane has taken its particles: the ring slot can be refilled
      if (lane == 0 && g + NS < ngroups) load_group(g + NS);
      // ---- B. gather (quadratics in d), velocity update, move
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
        // offsets from node c in cells: new offset tn, mid offset tm
        u[kb] = v[kb][0] * cells_per_v;
        tn[kb] = d[kb] + u[kb];
        tm[kb] = fma(R(0.5), u[kb], d[kb]);
        fast[kb] = valid[kb] && fast_bin && (fabs(tn[kb]) < R(1.5));
        if (STAG) fast[kb] = fast[kb] && (fabs(ts[kb]) < R(1.5));
      }
      // ---- C. deposit moments (all_fast: the common case, warp-uniform, no per-lane branches)
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) all_fast[kb] = __all_sync(0xffffffffu, fast[kb]);
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        if (all_fast[kb] || fast[kb]) {
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
        }
      }
      // ---- D. per block: destination, slot claim, retire the block stashed one group ago, stash the new one
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        // destination: 0 stay, 1 left, 2 right (3 = general path, -1 = no particle)
        const bool go_r = fast[kb] && tn[kb] >= R(0.5), go_l = fast[kb] && tn[kb] < R(-0.5);
        R dn = tn[kb];
        if (go_r) dn -= R(1);
        if (go_l) dn += R(1);
        const int kind = fast[kb] ? (go_r ? 2 : (go_l ? 1 : 0)) : (valid[kb] ? 3 : -1);
        // claim slots in the destination bins: one atomic per warp and destination, consumed one group later
        const unsigned m1 = __ballot_sync(0xffffffffu, go_l), m2 = __ballot_sync(0xffffffffu, go_r);
        const unsigned m0 = __ballot_sync(0xffffffffu, fast[kb]) & ~(m1 | m2);
        const unsigned my_cnt = __popc(lane == 0 ? m0 : (lane == 1 ? m1 : m2));
        unsigned claim = 0;
        if (lane < 3 && my_cnt) claim = atomicAdd(my_cursor, my_cnt);
        const unsigned rank = __popc((go_r ? m2 : (go_l ? m1 : m0)) & lt_mask);
        retire(kb);
        {
          R* st_ = my_stash + kb * kBlkElems;
          st_[0] = dn; st_[kBlk] = v[kb][0]; st_[2 * kBlk] = v[kb][1]; st_[3 * kBlk] = v[kb][2];
        }
        q_meta[kb] = fast[kb] ? (int)((rank << 2) | (unsigned)kind) : -1;
        q_claim[kb] = claim;
        if (!all_fast[kb]) {
          const unsigned ms = __ballot_sync(0xffffffffu, kind == 3);
          if (ms) {
            n_slow += __popc(ms);
            if (kind == 3) slow_tail(p, bd, dst, acc, s, node_pos(c, p) + d[kb] * p.dx, STAG ? vx_old[kb] : R(0), v[kb][0], v[kb][1], v[kb][2]);
          }
        }
      }
    }
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) retire(kb);  // the last group of the item

    // ---- flush: warp totals of the moments -> 19 node values -> atomics on the raw (L2-resident) grid
    if (n_slow < n) {
      R vals[18] = {r1, r2, rP, rN, y0, y1, y2, yP, yN, z0, z1, z2, zP, zN, a1, a2, aP, aN};
#pragma unroll
      for (int j = 0; j < 18; ++j) {
        const R tsum = warp_sum(vals[j]);
        if (lane == j) tot_s[warp][j] = tsum;
      }
      if (STAG) {
        R sv[5] = {s1, s2, sN, sZ, sP};
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const R tsum = warp_sum(sv[j]);
          if (lane == j) tot_s[warp][18 + j] = tsum;
        }
      }
      __syncwarp();
      if (STAG && lane >= 19 && lane < 25) {  // rho(x_n) on the faces c-3..c+2 (the truncated-power sums carry a factor 4)
        const R* tot = tot_s[warp];
        const R X0 = (R)(n - n_slow), X1 = tot[18], X2 = tot[19], Nm = R(0.25) * tot[20], Z = R(0.25) * tot[21], Pp = R(0.25) * tot[22];
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
        const R* tot = tot_s[warp];
        const R cnt = (R)(n - n_slow);
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

  // ---- particles that did not fit their bin last step: general path, one by one
  const int n_ov = min(hdr->ov_n[src], bd.ov_cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_ov; i += gridDim.x * blockDim.x) {
    const int b = bd.ov_bin[src][i];
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
