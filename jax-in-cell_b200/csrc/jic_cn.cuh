// Implicit Crank-Nicolson stepper (time_evolution_algorithm = 1): jaxincell/_algorithms.py:100-241.
//
// One reference step is a Picard fixed-point loop over
//   B^{n+1} = B^n - dt curl(E^{n+1/2})                      (Faraday with the average of E^n and the current guess)
//   sub-stepped particle push in E^{n+1/2}, B^{n+1/2}        (periodic S2 gather, Boris velocity update, midpoint move)
//   E^{n+1} = E^n + dt (c^2 curl B^{n+1/2} - (J - <J>)/eps0) (J = time average of the sub-step currents q v_mid S(x_stag))
// until |max(E_new - E_guess)| / (max|E_new| + 1e-12) <= tol or the iteration cap is hit (a data-dependent lax.while_loop).
//
// Here: `max_iter` x (k_cn_push -> [all-reduce] -> k_cn_fields) are enqueued per step, and the kernels of the iterations after
// convergence return at once on a device flag, so the whole loop stays inside a CUDA graph with no host round trip.
//   * k_cn_push: one thread per particle runs ALL sub-steps of one Picard iteration: its (x_n, v_n) are read once, the staggered
//     positions of the previous iteration (one real per sub-step) are read and replaced, (x_{n+1}, v_{n+1}) candidates are written
//     to the other half of a double buffer that the host swaps every step.  The gather and the current deposit of a sub-step share
//     their three weights (both live on the faces, _algorithms.py:110).  J and rho(x_{n+1}) go to the raw grid by atomics.
//   * k_cn_fields: single CTA: <J>, Ampere, the convergence test, then either the step outputs + the set-up of the next step,
//     or Faraday + the averaged tables for the next iteration.
// Reference quirks kept on purpose: the push always uses the non-relativistic Boris update and ignores external fields, the
// filter and field_solver; v_new is carried to the next sub-step WITHOUT the boundary flip that v_mid gets; the charges that enter
// J and rho are the step-start ones (an absorbed particle's charge comes back next step).
#pragma once
#include "jic_device.cuh"
#include "jic_kernels.cuh"

namespace jic {

struct CnControl {
  int converged;   // set by k_cn_fields when the Picard loop of the current step has ended
  int iter;        // Picard iterations executed in the current step
  int last_iters;  // ... in the last completed step
  int pad;
  long long total_iters;
};

template <typename R>
struct CnState {   // SoA particle state of one side of the double buffer
  R *x, *y, *z, *vx, *vy, *vz;
};

// periodic S2 stencil of _sources.py:10-40: nearest node k = round-half-even((x - start)/dx), nodes k-1,k,k+1 wrapped
template <typename R>
__device__ __forceinline__ void cn_stencil(R x, R start, const DevParams<R>& p, int idx[3], R w[3]) {
  const R xn = (x - start) / p.dx;
  const R kf = rint(xn);
  const int k = (int)kf;
  const R d = xn - kf;
  idx[0] = mod_pos(k - 1, p.G); idx[1] = mod_pos(k, p.G); idx[2] = mod_pos(k + 1, p.G);
  w[0] = R(0.5) * (R(0.5) - d) * (R(0.5) - d);
  w[1] = R(0.75) - d * d;
  w[2] = R(0.5) * (R(0.5) + d) * (R(0.5) + d);
}

// start-up of the CN carry (_simulation.py:216-220,237-240): positions stay x0, velocities are the post-BC ones of the
// half-step; rho0 (original charges) for the initial Gauss solve (_state_initialization.py:374)
template <typename R>
__global__ void __launch_bounds__(256) k_cn_start(const DevParams<R> p, const R* __restrict__ x0, const R* __restrict__ v0, CnState<R> s,
                                                  R* __restrict__ v_init, uint8_t* __restrict__ alive, R* __restrict__ acc) {
  const GlobalGrid<R> grid{acc};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.N; i += (long long)gridDim.x * blockDim.x) {
    const int sp = species_of(i, p);
    const R X0 = x0[3 * i];
    R v[3] = {v0[3 * i], v0[3 * i + 1], v0[3 * i + 2]};
    deposit_cloud(grid, make_cloud(X0, p), p.G, R(0), R(0), p.sp_q[sp] * p.inv_dx, false);
    R xp = X0 + p.half_dt * v[0];
    const int flag = bc_x(xp, p);
    if (flag == 1) v[0] = -v[0];
    if (flag == 2) { v[0] = v[1] = v[2] = R(0); }
    // a particle absorbed by the start-up half step keeps its position x0 but enters the CN carry with q = q/m = 0 for the
    // whole run (_simulation.py:217-220,237-240): remembered in a byte, because charges are per-species constants here
    alive[i] = flag != 2;
    s.x[i] = X0; s.y[i] = x0[3 * i + 1]; s.z[i] = x0[3 * i + 2];
    s.vx[i] = v[0]; s.vy[i] = v[1]; s.vz[i] = v[2];
    if (v_init) { v_init[3 * i] = v[0]; v_init[3 * i + 1] = v[1]; v_init[3 * i + 2] = v[2]; }
  }
}

// One Picard iteration of the particle part: _algorithms.py:148-188.  `it` is the iteration index of this launch.
// SHARED: persistent CTAs deposit into a CTA-private copy of the raw grid in shared memory (flushed once), like k_step.
template <typename R, bool SHARED>
__global__ void __launch_bounds__(256) k_cn_push(const DevParams<R> p, CnState<R> cur, CnState<R> nxt, R* __restrict__ stag, int n_sub, int it,
                                                 const double* __restrict__ Eavg, const double* __restrict__ Bavg, R* __restrict__ acc,
                                                 const uint8_t* __restrict__ alive, const CnControl* __restrict__ cn) {
  if (it > 0 && cn->converged) return;
  extern __shared__ __align__(16) unsigned char cn_smem_raw[];
  R* sacc = reinterpret_cast<R*>(cn_smem_raw);
  if (SHARED) {
    for (int k = threadIdx.x; k < p.G * kAccRow; k += blockDim.x) sacc[k] = R(0);
    __syncthreads();
  }
  R* const dep = SHARED ? sacc : acc;  // where this CTA's atomics go
  const R dtau = p.dt / R(n_sub), half_dtau = R(0.5) * dtau;
  const R e_start = p.g0 + p.half_dx, b_start = p.g0 - p.half_dx;  // :110-111
  const R w_sub = dtau / p.dt;                                       // J_iter = sum_s J_s dtau / dt  (:179,:190)
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.N; i += (long long)gridDim.x * blockDim.x) {
    const int sp = species_of(i, p);
    const bool live = alive[i] != 0;
    const R q = live ? p.sp_q[sp] : R(0), qm = live ? p.sp_qm[sp] : R(0);
    R pos[3] = {cur.x[i], cur.y[i], cur.z[i]};
    R vel[3] = {cur.vx[i], cur.vy[i], cur.vz[i]};
    const R x_n = pos[0];
    for (int s = 0; s < n_sub; ++s) {
      // staggered position of this sub-step from the previous iteration (all equal to x_n in the first one, :121)
      const R xs = it == 0 ? x_n : stag[(size_t)s * p.N + i];
      int ie[3], ib[3];
      R we[3], wb[3];
      cn_stencil(xs, e_start, p, ie, we);
      cn_stencil(xs, b_start, p, ib, wb);
      R E[3] = {0, 0, 0}, B[3] = {0, 0, 0};
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          E[c] += we[k] * (R)__ldg(Eavg + ie[k] * 3 + c);
          B[c] += wb[k] * (R)__ldg(Bavg + ib[k] * 3 + c);
        }
      R vnew[3] = {vel[0], vel[1], vel[2]};
      boris_velocity(vnew, E, B, qm, dtau);  // :157 (always the non-relativistic update, with the step-start q/m)
      R vmid[3] = {R(0.5) * (vel[0] + vnew[0]), R(0.5) * (vel[1] + vnew[1]), R(0.5) * (vel[2] + vnew[2])};
      pos[0] += vmid[0] * dtau; pos[1] += vmid[1] * dtau; pos[2] += vmid[2] * dtau;
      const int flag = bc_x(pos[0], p);  // :163-166 (the flip / zero applies to v_mid only)
      pos[1] = wrap_transverse(pos[1], p.Ly, p.half_Ly);
      pos[2] = wrap_transverse(pos[2], p.Lz, p.half_Lz);
      if (flag == 1) vmid[0] = -vmid[0];
      if (flag == 2) { vmid[0] = vmid[1] = vmid[2] = R(0); }
      R xst = pos[0] - half_dtau * vmid[0];  // :167-170
      bc_x(xst, p);
      stag[(size_t)s * p.N + i] = xst;
      // J_s = (q/dx) v_mid S(x_stag_prev) on the faces: same stencil as the E gather (:176-179)
      const R a = q * p.inv_dx * w_sub;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const R wk = we[k] * a;
        if (wk != R(0)) {
          atomicAdd(dep + ie[k] * kAccRow + 0, wk * vmid[0]);
          atomicAdd(dep + ie[k] * kAccRow + 1, wk * vmid[1]);
          atomicAdd(dep + ie[k] * kAccRow + 2, wk * vmid[2]);
        }
      }
      vel[0] = vnew[0]; vel[1] = vnew[1]; vel[2] = vnew[2];
    }
    nxt.x[i] = pos[0]; nxt.y[i] = pos[1]; nxt.z[i] = pos[2];
    nxt.vx[i] = vel[0]; nxt.vy[i] = vel[1]; nxt.vz[i] = vel[2];
    // rho(x_{n+1}) for the step output (:236-238): ordinary S2 cloud with the particle-BC fold, step-start charge, no filter
    const GlobalGrid<R> grid{dep};
    deposit_cloud(grid, make_cloud(pos[0], p), p.G, R(0), R(0), q * p.inv_dx, false);
  }
  if (SHARED) {
    __syncthreads();
    for (int k = threadIdx.x; k < p.G * kAccRow; k += blockDim.x) {
      const R v = sacc[k];
      if (v != R(0)) atomicAdd(acc + k, v);
    }
  }
}

// histories of the particles for the step that just ended (only launched when the caller asked for them)
template <typename R>
__global__ void __launch_bounds__(256) k_cn_record(const DevParams<R> p, CnState<R> s, const RunControl* __restrict__ ctl) {
  R* x_hist = (R*)ctl->hist[4];
  R* v_hist = (R*)ctl->hist[5];
  if (!x_hist && !v_hist) return;
  const long long row = ctl->hist_row - 1;  // k_cn_fields has already advanced it
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.N; i += (long long)gridDim.x * blockDim.x) {
    if (x_hist) { R* o = x_hist + ((size_t)row * p.N + i) * 3; o[0] = s.x[i]; o[1] = s.y[i]; o[2] = s.z[i]; }
    if (v_hist) { R* o = v_hist + ((size_t)row * p.N + i) * 3; o[0] = s.vx[i]; o[1] = s.vy[i]; o[2] = s.vz[i]; }
  }
}

template <typename R>
struct CnFieldArgs {
  int G, fbl, fbr, it, max_iter, prepare_only;
  double dx, dt, tol;
  R* acc;                  // raw (G,4): J_iter (already weighted by dtau/dt) and rho(x_{n+1}); consumed and zeroed
  double *En, *Bn;         // fields at the start of the step (G,3); replaced when the step ends
  double *Eg, *Bnext;      // current guess of E^{n+1}; B^{n+1} of the current iteration
  double *Eavg, *Bavg;     // what the push gathers
  double* EB;              // optional (sorted push): the same packed per node, (E_avg[k], B_avg[k + 1]) -- see cn_prepare
  double *J, *rho;         // outputs of the step
  CnControl* cn;
  RunControl* ctl;
};

__device__ __forceinline__ double block_reduce(double v, bool take_max, double* sm /* >= 33 */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int o = 16; o; o >>= 1) { const double u = __shfl_xor_sync(0xffffffffu, v, o); v = take_max ? fmax(v, u) : v + u; }
  __syncthreads();
  if (lane == 0) sm[w] = v;
  __syncthreads();
  if (w == 0) {
    v = lane < nw ? sm[lane] : (take_max ? -INFINITY : 0.0);
    for (int o = 16; o; o >>= 1) { const double u = __shfl_xor_sync(0xffffffffu, v, o); v = take_max ? fmax(v, u) : v + u; }
    if (lane == 0) sm[32] = v;
  }
  __syncthreads();
  return sm[32];
}

// Faraday with the averaged E and the tables the push gathers: _algorithms.py:133-142
// EB (optional): row k = (E_avg[k], B_avg[(k + 1) mod G]).  The faces of B sit one node left of those of E (_algorithms.py:110-111), so
// the B stencil of a particle is its E stencil shifted by one node with the same weights: one packed row per stencil node serves both.
__device__ __forceinline__ void cn_prepare(const double* En, const double* Bn, const double* Eg, double* Bnext, double* Eavg, double* Bavg,
                                           int G, int fbl, double dx, double dt, double* EB = nullptr) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int k = tid; k < G * 3; k += nt) Eavg[k] = 0.5 * (En[k] + Eg[k]);
  __syncthreads();
  double gl[3];
  ghost_E_left(Eavg, Bn, G, fbl, gl);  // curlE(E_avg, B_field, ...): the ghost row is built from (E_avg, B^n)
  for (int i = tid; i < G; i += nt) {
    const double ey_m = i ? Eavg[(i - 1) * 3 + 1] : gl[1], ez_m = i ? Eavg[(i - 1) * 3 + 2] : gl[2];
    const double dFz = (Eavg[i * 3 + 2] - ez_m) / dx, dFy = (Eavg[i * 3 + 1] - ey_m) / dx;
    const double bn0 = Bn[i * 3], bn1 = Bn[i * 3 + 1], bn2 = Bn[i * 3 + 2];
    const double b0 = bn0 - dt * 0.0, b1 = bn1 - dt * (-dFz), b2 = bn2 - dt * dFy;
    Bnext[i * 3] = b0; Bnext[i * 3 + 1] = b1; Bnext[i * 3 + 2] = b2;
    Bavg[i * 3] = 0.5 * (bn0 + b0); Bavg[i * 3 + 1] = 0.5 * (bn1 + b1); Bavg[i * 3 + 2] = 0.5 * (bn2 + b2);
  }
  __syncthreads();
  if (EB) {
    for (int i = tid; i < G; i += nt) {
      const int ib = i + 1 < G ? i + 1 : 0;
#pragma unroll
      for (int c = 0; c < 3; ++c) { EB[i * 6 + c] = Eavg[i * 3 + c]; EB[i * 6 + 3 + c] = Bavg[ib * 3 + c]; }
    }
  }
}

template <typename R>
__global__ void __launch_bounds__(1024) k_cn_fields(const CnFieldArgs<R> a) {
  __shared__ double sm[40];
  const int G = a.G, tid = threadIdx.x, nt = blockDim.x;
  if (a.prepare_only) {  // after the initial Gauss solve: guess = E^n, first Faraday
    for (int k = tid; k < G * 3; k += nt) a.Eg[k] = a.En[k];
    if (tid == 0) { a.cn->converged = 0; a.cn->iter = 0; }
    __syncthreads();
    cn_prepare(a.En, a.Bn, a.Eg, a.Bnext, a.Eavg, a.Bavg, G, a.fbl, a.dx, a.dt, a.EB);
    return;
  }
  if (a.it > 0 && a.cn->converged) return;
  // <J> per component over the grid (:191), J_iter
  double s0 = 0, s1 = 0, s2 = 0;
  for (int i = tid; i < G; i += nt) {
    const double j0 = (double)a.acc[i * kAccRow], j1 = (double)a.acc[i * kAccRow + 1], j2 = (double)a.acc[i * kAccRow + 2];
    a.J[i * 3] = j0; a.J[i * 3 + 1] = j1; a.J[i * 3 + 2] = j2;
    a.rho[i] = (double)a.acc[i * kAccRow + 3];
    s0 += j0; s1 += j1; s2 += j2;
  }
  const double m0 = block_reduce(s0, false, sm) / G, m1 = block_reduce(s1, false, sm) / G, m2 = block_reduce(s2, false, sm) / G;
  for (int k = tid; k < G * kAccRow; k += nt) a.acc[k] = R(0);
  // Ampere (:196-198): E_calc = E^n + dt (c^2 curl B_avg - (J - <J>)/eps0); curlB's ghost row from (B_avg, E^n).  E_calc -> Eavg scratch
  double gr[3];
  ghost_B_right(a.Bavg, a.En, G, a.fbr, gr);
  double dmax = -INFINITY, emax = 0.0;
  __syncthreads();
  for (int i = tid; i < G; i += nt) {
    const double by_p = (i + 1 < G) ? a.Bavg[(i + 1) * 3 + 1] : gr[1], bz_p = (i + 1 < G) ? a.Bavg[(i + 1) * 3 + 2] : gr[2];
    const double dFz = (bz_p - a.Bavg[i * 3 + 2]) / a.dx, dFy = (by_p - a.Bavg[i * 3 + 1]) / a.dx;
    const double e0 = a.En[i * 3] + a.dt * ((kC * kC) * 0.0 - (1 / kEps0) * (a.J[i * 3] - m0));
    const double e1 = a.En[i * 3 + 1] + a.dt * ((kC * kC) * (-dFz) - (1 / kEps0) * (a.J[i * 3 + 1] - m1));
    const double e2 = a.En[i * 3 + 2] + a.dt * ((kC * kC) * dFy - (1 / kEps0) * (a.J[i * 3 + 2] - m2));
    dmax = fmax(dmax, fmax(e0 - a.Eg[i * 3], fmax(e1 - a.Eg[i * 3 + 1], e2 - a.Eg[i * 3 + 2])));
    emax = fmax(emax, fmax(fabs(e0), fmax(fabs(e1), fabs(e2))));
    a.Eavg[i * 3] = e0; a.Eavg[i * 3 + 1] = e1; a.Eavg[i * 3 + 2] = e2;
  }
  dmax = block_reduce(dmax, true, sm);
  emax = block_reduce(emax, true, sm);
  const double delta = fabs(dmax) / (emax + 1e-12);  // :224 (abs of the max, as written)
  const int iter = a.cn->iter + 1;
  const bool more = (delta > a.tol) && (iter < a.max_iter);  // :213-215
  __syncthreads();
  for (int k = tid; k < G * 3; k += nt) a.Eg[k] = a.Eavg[k];
  __syncthreads();
  if (more) {
    if (tid == 0) { a.cn->iter = iter; a.cn->converged = 0; }
    cn_prepare(a.En, a.Bn, a.Eg, a.Bnext, a.Eavg, a.Bavg, G, a.fbl, a.dx, a.dt, a.EB);
    return;
  }
  // ---- the step ends: E^{n+1} = last E_calc, B^{n+1} = the B_next this iteration pushed with (:229-241)
  const long long row = a.ctl->hist_row;
  R* hE = (R*)a.ctl->hist[0]; R* hB = (R*)a.ctl->hist[1]; R* hJ = (R*)a.ctl->hist[2]; R* hrho = (R*)a.ctl->hist[3];
  for (int k = tid; k < G * 3; k += nt) {
    const double e = a.Eg[k], b = a.Bnext[k];
    a.En[k] = e; a.Bn[k] = b;
    if (hE) hE[(size_t)row * G * 3 + k] = (R)e;
    if (hB) hB[(size_t)row * G * 3 + k] = (R)b;
    if (hJ) hJ[(size_t)row * G * 3 + k] = (R)a.J[k];
  }
  if (hrho) for (int i = tid; i < G; i += nt) hrho[(size_t)row * G + i] = (R)a.rho[i];
  __syncthreads();
  if (tid == 0) {
    a.cn->converged = 1; a.cn->last_iters = iter; a.cn->total_iters += iter; a.cn->iter = 0;
    a.ctl->hist_row = row + 1; a.ctl->step += 1;
  }
  // first Faraday of the next step (guess = the new E^n)
  cn_prepare(a.En, a.Bn, a.Eg, a.Bnext, a.Eavg, a.Bavg, G, a.fbl, a.dx, a.dt, a.EB);
}

}  // namespace jic
