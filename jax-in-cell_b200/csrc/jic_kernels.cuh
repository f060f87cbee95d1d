// Kernels of the INDEXED engine (particle p stays in slot p) and the single-CTA field kernel shared by both engines.
#pragma once
#include "jic_device.cuh"

namespace jic {

// ---------------------------------------------------------------------------------------------------------
// K0a  leap-frog start-up + initial deposits.
//   jaxincell/_simulation.py:217-225   x_{+1/2}, v, q <- BC(x0 + dt/2 v0);  x_{-1/2} <- BCpos(x0 - dt/2 v) with post-BC v
//   jaxincell/_state_initialization.py:374   rho0 from x0 with the ORIGINAL charges
//   jaxincell/_algorithms.py:29-32    J^0 from (x_{-1/2}, x0, x_{+1/2}, v, q) with the post-BC charges
// ---------------------------------------------------------------------------------------------------------
//   x0, v0 hold particles [i0, i0 + n) of the run (the whole run, or one chunk of a pipelined host upload).
template <typename R>
__global__ void __launch_bounds__(256) k_start(const DevParams<R> p, const R* __restrict__ x0, const R* __restrict__ v0, long long i0, long long n,
                                               R* __restrict__ xh, R* __restrict__ yh, R* __restrict__ zh, R* __restrict__ vx,
                                               R* __restrict__ vy, R* __restrict__ vz, R* __restrict__ v_init, R* __restrict__ acc) {
  const GlobalGrid<R> grid{acc};
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
    const long long i = i0 + j;
    const int s = species_of(i, p);
    const R q = p.sp_q[s];
    const R X0 = x0[3 * j], Y0 = x0[3 * j + 1], Z0 = x0[3 * j + 2];
    R v[3] = {v0[3 * j], v0[3 * j + 1], v0[3 * j + 2]};
    // rho0 (original charge, raw initial position)
    const Cloud<R> c0 = make_cloud(X0, p);
    deposit_cloud(grid, c0, p.G, R(0), R(0), q * p.inv_dx, false);
    // x_{+1/2}
    R xp = X0 + p.half_dt * v[0];
    const int flag = bc_x(xp, p);
    R qj = q;
    if (flag == 1) v[0] = -v[0];
    if (flag == 2) { v[0] = v[1] = v[2] = R(0); qj = R(0); }
    // x_{-1/2} with the post-BC velocity
    R xm = X0 - p.half_dt * v[0];
    bc_x(xm, p);
    if (qj != R(0)) {
      const Cloud<R> cm = make_cloud(xm, p), cp = make_cloud(xp, p);
      deposit_jx_startup(grid, xm, cm, cp, qj / p.dt, p);
      // J_y,z = rho(x_0) v_{y,z}: re-use the x0 cloud (same position, post-BC charge)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int k = c0.c + j - 1;
        if (k >= 0 && k < p.G) { grid.add(k, 1, c0.w[j] * qj * p.inv_dx * v[1]); grid.add(k, 2, c0.w[j] * qj * p.inv_dx * v[2]); }
      }
      if (c0.first != R(0)) { grid.add(0, 1, c0.first * qj * p.inv_dx * v[1]); grid.add(0, 2, c0.first * qj * p.inv_dx * v[2]); }
      if (c0.last != R(0)) { grid.add(p.G - 1, 1, c0.last * qj * p.inv_dx * v[1]); grid.add(p.G - 1, 2, c0.last * qj * p.inv_dx * v[2]); }
    }
    xh[i] = xp;
    vx[i] = v[0]; vy[i] = v[1]; vz[i] = v[2];
    if (p.track_yz) {
      yh[i] = wrap_transverse(Y0 + p.half_dt * v0[3 * j + 1], p.Ly, p.half_Ly);
      zh[i] = wrap_transverse(Z0 + p.half_dt * v0[3 * j + 2], p.Lz, p.half_Lz);
    }
    if (v_init) { v_init[3 * i] = v[0]; v_init[3 * i + 1] = v[1]; v_init[3 * i + 2] = v[2]; }
  }
}

// ---------------------------------------------------------------------------------------------------------
// K1  fused gather -> Boris push -> particle BC -> x_{n+1} -> deposit (J_x, J_y, J_z, rho), one pass over the
//     particle arrays: jaxincell/_algorithms.py:40-66 and :90-92 (the first deposit of the next reference step,
//     :29-32, is this same deposit -- SURVEY.md section 0).
//     SHARED = CTA-private copy of the raw grid in shared memory (persistent CTAs, flushed once).
// ---------------------------------------------------------------------------------------------------------
template <typename R, bool SHARED>
__global__ void __launch_bounds__(256) k_step(const DevParams<R> p, R* __restrict__ xh, R* __restrict__ yh, R* __restrict__ zh,
                                              R* __restrict__ vx, R* __restrict__ vy, R* __restrict__ vz, const R* __restrict__ F,
                                              R* __restrict__ acc, const RunControl* __restrict__ ctl) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  R* sacc = reinterpret_cast<R*>(smem_raw);
  if (SHARED) {
    for (int k = threadIdx.x; k < p.G * kAccRow; k += blockDim.x) sacc[k] = R(0);
    __syncthreads();
  }
  R* x_hist = (R*)ctl->hist[4];
  R* v_hist = (R*)ctl->hist[5];
  const long long row = (x_hist || v_hist) ? ctl->hist_row : 0;
  R* xrow = x_hist ? x_hist + (size_t)row * 3 * p.N : nullptr;
  R* vrow = v_hist ? v_hist + (size_t)row * 3 * p.N : nullptr;

  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.N; i += (long long)gridDim.x * blockDim.x) {
    const R x_old = xh[i];
    R v[3] = {vx[i], vy[i], vz[i]};
    // absorbed particles are parked outside the box with v = 0, q = 0 (_boundary_conditions.py:40,51,64,78): nothing moves
    const bool dead = (x_old < -p.half_L) || (x_old > p.half_L);
    R x_new = x_old, x_mid = x_old, y_mid = R(0), z_mid = R(0);
    R vpre_y = R(0), vpre_z = R(0);  // transverse velocity the pusher advanced y,z with (before an absorption zeroes it)
    if (!dead) {
      const int s = species_of(i, p);
      R E[3], B[3];
      const R vx_old = v[0];
      gather_fields(F, x_old, p, E, B);
      if (p.relativistic) boris_velocity_relativistic(v, E, B, p.sp_q[s], p.sp_m[s], p.dt);
      else boris_velocity(v, E, B, p.sp_qm[s], p.dt);
      vpre_y = v[1]; vpre_z = v[2];
      x_new = x_old + p.dt * v[0];
      const int flag = bc_x(x_new, p);
      R q = p.sp_q[s];
      if (flag == 1) v[0] = -v[0];
      if (flag == 2) { v[0] = v[1] = v[2] = R(0); q = R(0); }
      x_mid = x_new - p.half_dt * v[0];
      bc_x(x_mid, p);
      if (q != R(0)) {
        const Cloud<R> c_old = make_cloud(x_old, p), c_new = make_cloud(x_new, p), c_mid = make_cloud(x_mid, p);
        const R a = q * p.inv_dx;
        if (SHARED) {
          const SharedGrid<R> g{sacc};
          deposit_jx(g, x_old, c_old, c_new, q / p.dt, p);
          deposit_cloud(g, c_mid, p.G, a * v[1], a * v[2], a, true);
        } else {
          const GlobalGrid<R> g{acc};
          deposit_jx(g, x_old, c_old, c_new, q / p.dt, p);
          deposit_cloud(g, c_mid, p.G, a * v[1], a * v[2], a, true);
        }
        if (p.stag) {
          // rho on the faces at x_n = BCpos(x_{n+1/2} - dt/2 v_n), the position the reference still holds in `positions` at
          // _algorithms.py:70 (bit-identical recomputation of _algorithms.py:60-61 of the previous step), post-BC charge
          R x_n = x_old - p.half_dt * vx_old;
          bc_x(x_n, p);
          deposit_faces(GlobalGrid<R>{acc, p.G}, make_cloud_faces(x_n, p), p.G, a);
        }
      }
      xh[i] = x_new;
      vx[i] = v[0]; vy[i] = v[1]; vz[i] = v[2];
    }
    if (p.track_yz) {
      // y,z: advanced with the pushed velocity (_particles.py:125), wrapped by the BC (_boundary_conditions.py:28-29),
      // then x_{n+1} = BCpos(x_{n+3/2} - dt/2 v_{n+1}) with the post-BC velocity (_algorithms.py:60-61)
      const R y = wrap_transverse(yh[i] + p.dt * vpre_y, p.Ly, p.half_Ly);
      const R z = wrap_transverse(zh[i] + p.dt * vpre_z, p.Lz, p.half_Lz);
      yh[i] = y; zh[i] = z;
      y_mid = wrap_transverse(y - p.half_dt * v[1], p.Ly, p.half_Ly);
      z_mid = wrap_transverse(z - p.half_dt * v[2], p.Lz, p.half_Lz);
    }
    if (xrow) { xrow[3 * i] = x_mid; xrow[3 * i + 1] = y_mid; xrow[3 * i + 2] = z_mid; }
    if (vrow) { vrow[3 * i] = v[0]; vrow[3 * i + 1] = v[1]; vrow[3 * i + 2] = v[2]; }
  }

  if (SHARED) {
    __syncthreads();
    for (int k = threadIdx.x; k < p.G * kAccRow; k += blockDim.x) {
      const R v = sacc[k];
      if (v != R(0)) atomicAdd(acc + k, v);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// K2  single-CTA field kernel: digital filter of the raw [Jx,Jy,Jz,rho] grid (jaxincell/_filters.py:52-153),
//     second Maxwell half step of step n (B then E, _fields.py:185-193), emission of the step outputs
//     (_algorithms.py:93), first Maxwell half step of step n+1 (E then B, _fields.py:175-183) and the padded
//     total-field table for the next gather (_algorithms.py:36-37 + _boundary_conditions.py:233-245).
//     In `init` mode it instead turns rho0 into E_x (prefix sum == the dense solve of _fields.py:76-81).
//     Field arithmetic is always fp64; only the raw grid, the gather table and the histories are `R`.
// ---------------------------------------------------------------------------------------------------------
template <typename R>
struct FieldArgs {
  int G, fbl, fbr, passes, n_strides, init;
  int strides[JIC_MAX_STRIDES];
  double alpha, dx, dt;
  R* acc;              // raw grid (G,4), consumed and zeroed
  double *E, *B;       // (G,3) leap-frog state
  double *E_int, *B_int;  // copies at integer time (what the step emits)
  double *J, *rho;     // filtered (G,3), (G)
  const double *extE, *extB;
  R* F;                // padded gather table (G+3, 8)
  double *s0, *s1;     // filter scratch (G,4) each
  double *E0, *B0;     // initial fields (init mode)
  int smem_comps;      // components filtered at a time in shared memory (0 = global scratch)
  const double* ExC;   // field_solver != 0: E_x of this step from k_gauss (G), else null; the raw grid then has a fifth component
  int record;          // write this step's outputs to the history buffers named in ctl->hist
  RunControl* ctl;
};

// neighbour j+s of a component stored contiguously (y[0..G)): jaxincell/_filters.py:9-50 (_shift_with_bc_1d)
__device__ __forceinline__ double filt_neighbour(const double* y, int j, int s, int G, bool periodic, int fbl, int fbr) {
  int k = j + s;
  if (periodic) {
    if (k < 0) { k += G; if (k < 0) k = mod_pos(k, G); }
    else if (k >= G) { k -= G; if (k >= G) k = mod_pos(k, G); }
    return y[k];
  }
  if (k < 0) return fbl == JIC_BC_ABSORBING ? 0.0 : y[0];
  if (k >= G) return fbr == JIC_BC_ABSORBING ? 0.0 : y[G - 1];
  return y[k];
}

__device__ __forceinline__ void ghost_E_left(const double* E, const double* B, int G, int fbl, double g[3]) {
  if (fbl == JIC_BC_PERIODIC) { g[0] = E[(G - 1) * 3]; g[1] = E[(G - 1) * 3 + 1]; g[2] = E[(G - 1) * 3 + 2]; }
  else if (fbl == JIC_BC_REFLECTIVE) { g[0] = E[0]; g[1] = E[1]; g[2] = E[2]; }
  else { g[0] = 0.0; g[1] = -2 * kC * B[2] - E[1]; g[2] = 2 * kC * B[1] - E[2]; }
}

__device__ __forceinline__ void ghost_B_right(const double* B, const double* E, int G, int fbr, double g[3]) {
  const int l = (G - 1) * 3;
  if (fbr == JIC_BC_PERIODIC) { g[0] = B[0]; g[1] = B[1]; g[2] = B[2]; }
  else if (fbr == JIC_BC_REFLECTIVE) { g[0] = B[l]; g[1] = B[l + 1]; g[2] = B[l + 2]; }
  else { g[0] = 0.0; g[1] = -(2 / kC) * E[l + 2] - B[l + 1]; g[2] = (2 / kC) * E[l + 1] - B[l + 2]; }
}

// B -= h curl E   (backward difference, left ghost; _fields.py:102-111, :182/:189)
__device__ __forceinline__ void faraday(double* E, double* B, int G, int fbl, double dx, double h) {
  double gl[3];
  ghost_E_left(E, B, G, fbl, gl);
  __syncthreads();  // every thread has read the ghost inputs before B[0] changes
  for (int i = threadIdx.x; i < G; i += blockDim.x) {
    const double ey_m = i ? E[(i - 1) * 3 + 1] : gl[1], ez_m = i ? E[(i - 1) * 3 + 2] : gl[2];
    const double dFz = (E[i * 3 + 2] - ez_m) / dx, dFy = (E[i * 3 + 1] - ey_m) / dx;
    B[i * 3 + 1] -= h * (-dFz);
    B[i * 3 + 2] -= h * dFy;
  }
  __syncthreads();
}

// E += h (c^2 curl B - J/eps0)   (forward difference, right ghost; _fields.py:132-144, :179/:192)
__device__ __forceinline__ void ampere(double* E, double* B, const double* J, int G, int fbr, double dx, double h) {
  double gr[3];
  ghost_B_right(B, E, G, fbr, gr);
  __syncthreads();
  for (int i = threadIdx.x; i < G; i += blockDim.x) {
    const double by_p = (i + 1 < G) ? B[(i + 1) * 3 + 1] : gr[1], bz_p = (i + 1 < G) ? B[(i + 1) * 3 + 2] : gr[2];
    const double dFz = (bz_p - B[i * 3 + 2]) / dx, dFy = (by_p - B[i * 3 + 1]) / dx;
    E[i * 3 + 0] += h * ((kC * kC) * 0.0 - (J[i * 3 + 0] / kEps0));
    E[i * 3 + 1] += h * ((kC * kC) * (-dFz) - (J[i * 3 + 1] / kEps0));
    E[i * 3 + 2] += h * ((kC * kC) * dFy - (J[i * 3 + 2] / kEps0));
  }
  __syncthreads();
}

// Digital filter of `cg` components held contiguously in cur[c * G + j] (double-buffered with nxt), all strides and sweeps:
// jaxincell/_filters.py:52-153.  Returns the buffer holding the result.  Called with shared-memory or global pointers; kept
// as a separate function so that the shared-memory instantiation compiles to 32-bit LDS/STS addressing.
template <typename Ptr>
__device__ __forceinline__ Ptr filter_components(Ptr cur, Ptr nxt, int cg, int G, int passes, double alpha, int n_strides, const int* strides,
                                                 int fbl, int fbr) {
  const int tid = threadIdx.x, nt = blockDim.x;
  if (passes <= 0) return cur;
  const bool periodic = (fbl == JIC_BC_PERIODIC) && (fbr == JIC_BC_PERIODIC);
  const int p_cl = passes < 17 ? passes : 17;
  const int n_reg = (passes - 1) < 16 ? (passes - 1) : 16;
  const double comp_alpha = p_cl - alpha * (p_cl - 1);
  const double zl = fbl == JIC_BC_ABSORBING ? 0.0 : 1.0, zr = fbr == JIC_BC_ABSORBING ? 0.0 : 1.0;
  for (int si = 0; si < n_strides; ++si) {
    const int s = strides[si];
    for (int sweep = 0; sweep <= n_reg; ++sweep) {
      const double al = sweep < n_reg ? alpha : comp_alpha;
      const double co = (1 - al) * 0.5;
      if (periodic && s < G) {
        for (int c = 0; c < cg; ++c) {
          const int o = c * G;
          for (int j = tid; j < G; j += nt) {
            int jm = j - s, jp = j + s;
            jm += jm < 0 ? G : 0;
            jp -= jp >= G ? G : 0;
            nxt[o + j] = al * cur[o + j] + co * (cur[o + jm] + cur[o + jp]);
          }
        }
      } else if (periodic) {
        for (int c = 0; c < cg; ++c) {
          const int o = c * G;
          for (int j = tid; j < G; j += nt) nxt[o + j] = al * cur[o + j] + co * (cur[o + mod_pos(j - s, G)] + cur[o + mod_pos(j + s, G)]);
        }
      } else {  // index clamp, and zero where the overrun side is absorbing (_filters.py:26-50)
        for (int c = 0; c < cg; ++c) {
          const int o = c * G;
          for (int j = tid; j < G; j += nt) {
            const int jm = j - s, jp = j + s;
            const double l = jm < 0 ? zl * cur[o] : cur[o + jm];
            const double r = jp >= G ? zr * cur[o + G - 1] : cur[o + jp];
            nxt[o + j] = al * cur[o + j] + co * (l + r);
          }
        }
      }
      __syncthreads();
      Ptr t = cur; cur = nxt; nxt = t;
    }
  }
  return cur;
}

template <typename R>
__global__ void __launch_bounds__(1024) k_fields(const FieldArgs<R> a) {
  extern __shared__ __align__(16) double fsm[];  // 2 * smem_comps * G doubles (0 = filter through the global scratch s0/s1)
  const int G = a.G, tid = threadIdx.x, nt = blockDim.x;
  // 1.+2. digital filter of the raw grid, `cg` components at a time
  {
    const int cg = a.smem_comps > 0 ? a.smem_comps : kAccRow;
    for (int c0 = 0; c0 < kAccRow; c0 += cg) {
      if (a.smem_comps > 0) {
        double* cur = fsm;
        for (int c = 0; c < cg; ++c)
          for (int j = tid; j < G; j += nt) cur[c * G + j] = (double)a.acc[j * kAccRow + c0 + c];
        __syncthreads();
        const double* res = filter_components<double*>(cur, fsm + cg * G, cg, G, a.passes, a.alpha, a.n_strides, a.strides, a.fbl, a.fbr);
        for (int c = 0; c < cg; ++c) {
          const int comp = c0 + c;
          for (int j = tid; j < G; j += nt) {
            if (comp < 3) a.J[j * 3 + comp] = res[c * G + j];
            else a.rho[j] = res[c * G + j];
          }
        }
      } else {
        double* cur = a.s0;
        for (int c = 0; c < cg; ++c)
          for (int j = tid; j < G; j += nt) cur[c * G + j] = (double)a.acc[j * kAccRow + c0 + c];
        __syncthreads();
        const double* res = filter_components<double*>(cur, a.s1, cg, G, a.passes, a.alpha, a.n_strides, a.strides, a.fbl, a.fbr);
        for (int c = 0; c < cg; ++c) {
          const int comp = c0 + c;
          for (int j = tid; j < G; j += nt) {
            if (comp < 3) a.J[j * 3 + comp] = res[c * G + j];
            else a.rho[j] = res[c * G + j];
          }
        }
      }
      __syncthreads();
    }
    // the raw grid is consumed: zero it for the next step (with the face component behind it when it is in use)
    for (int k = tid; k < G * (kAccRow + (a.ExC ? 1 : 0)); k += nt) a.acc[k] = R(0);
  }
  const double h = a.dt / 2;
  if (a.init) {
    // 3a. E_x = (dx/eps0) cumsum(rho0); the reference's forward substitution is this same sequential sum
    if (tid == 0) {
      double run = 0.0;
      for (int i = 0; i < G; ++i) { run += a.rho[i]; a.E[i * 3] = (a.dx / kEps0) * run; }
    }
    for (int i = tid; i < G; i += nt) {
      a.E[i * 3 + 1] = 0.0; a.E[i * 3 + 2] = 0.0;
      a.B[i * 3] = 0.0; a.B[i * 3 + 1] = 0.0; a.B[i * 3 + 2] = 0.0;
    }
    __syncthreads();
    for (int k = tid; k < G * 3; k += nt) { a.E0[k] = a.E[k]; a.B0[k] = a.B[k]; a.E_int[k] = a.E[k]; a.B_int[k] = a.B[k]; }
    __syncthreads();
  } else {
    // 3b. second half step of step n: B then E (_fields.py:185-193)
    faraday(a.E, a.B, G, a.fbl, a.dx, h);
    ampere(a.E, a.B, a.J, G, a.fbr, a.dx, h);
    if (a.ExC) {  // _algorithms.py:78: E_x is replaced by the Gauss / Poisson solve
      for (int i = tid; i < G; i += nt) a.E[i * 3] = a.ExC[i];
      __syncthreads();
    }
    const long long row = a.ctl->hist_row;
    R* hE = a.record ? (R*)a.ctl->hist[0] : nullptr;
    R* hB = a.record ? (R*)a.ctl->hist[1] : nullptr;
    R* hJ = a.record ? (R*)a.ctl->hist[2] : nullptr;
    R* hrho = a.record ? (R*)a.ctl->hist[3] : nullptr;
    for (int k = tid; k < G * 3; k += nt) {
      a.E_int[k] = a.E[k]; a.B_int[k] = a.B[k];
      if (hE) hE[(size_t)row * G * 3 + k] = (R)a.E[k];
      if (hB) hB[(size_t)row * G * 3 + k] = (R)a.B[k];
      if (hJ) hJ[(size_t)row * G * 3 + k] = (R)a.J[k];
    }
    if (hrho) for (int i = tid; i < G; i += nt) hrho[(size_t)row * G + i] = (R)a.rho[i];
    __syncthreads();
    if (tid == 0) { a.ctl->hist_row = row + 1; a.ctl->step += 1; }
  }
  // 4. first half step of the next step: E then B (_fields.py:175-183)
  ampere(a.E, a.B, a.J, G, a.fbr, a.dx, h);
  faraday(a.E, a.B, G, a.fbl, a.dx, h);
  // 5. padded total fields for the gather: rows [L2, L1, f_0..f_{G-1}, R]
  for (int r = tid; r < G + 3; r += nt) {
    int src;  // source cell, -1 = zeros
    if (r >= 2 && r < G + 2) src = r - 2;
    else if (r < 2) src = a.fbl == JIC_BC_PERIODIC ? (G - 2 + r) : a.fbl == JIC_BC_REFLECTIVE ? (1 - r) : -1;
    else src = a.fbr == JIC_BC_PERIODIC ? 0 : a.fbr == JIC_BC_REFLECTIVE ? (G - 1) : -1;
    if (src >= G) src = G - 1;   // G == 1 corner
    if (src < -1) src = 0;
    R* f = a.F + (size_t)r * kFieldRow;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      f[c] = src < 0 ? R(0) : (R)(a.E[src * 3 + c] + a.extE[src * 3 + c]);
      f[3 + c] = src < 0 ? R(0) : (R)(a.B[src * 3 + c] + a.extB[src * 3 + c]);
    }
    f[6] = R(0); f[7] = R(0);
  }
}

// ---------------------------------------------------------------------------------------------------------
// K2m  the field kernel on NC CTAs (used when the grid is large enough; same arithmetic as k_fields, non-init mode).
//   The filter and the Maxwell half steps are local stencils, so every CTA owns a slice of S nodes and recomputes a halo of
//   H = (total filter reach) + 2 nodes on each side from the raw grid: no inter-CTA communication, no grid-wide barrier.
//   Because CTAs read each other's slices (halos), the state is ping-ponged: this step reads acc_cur / E_r / B_r and writes
//   E_w / B_w, and zeroes its slice of acc_next (the buffer the NEXT push deposits into; it was consumed one step ago).
//   Non-periodic boundaries are handled by the edge CTAs with the clamp / zero rule of _filters.py:26-50 and the ghost-cell
//   formulas of _boundary_conditions.py:148-207; S >= H + 2 is guaranteed by the host so only they see the domain edge.
// ---------------------------------------------------------------------------------------------------------
template <typename R>
struct FieldArgsMC {
  int G, fbl, fbr, passes, n_strides;
  int strides[JIC_MAX_STRIDES];
  double alpha, dx, dt;
  int S, H, NC;
  const R* acc_cur; R* acc_next;
  const double* ExC;   // field_solver != 0: E_x of this step from k_gauss (G), else null
  const double *E_r, *B_r;
  double *E_w, *B_w, *E_int, *B_int, *J, *rho;
  const double *extE, *extB;
  R* F;
  int record;
  RunControl* ctl;
  unsigned* done;  // CTAs finished (the last one advances ctl->step / hist_row and resets it)
  // N GPUs, fused reduction: instead of an NCCL all-reduce in front of this kernel, every rank reads the raw
  // grids of all ranks straight from their memory over NVLink (peer mappings of the same buffer) while it loads its window, and
  // sums them in rank order -- the same order on every rank, so E and B stay bit-identical everywhere.
  int world, rank;                         // world <= 1: not fused
  const R* peer_acc[JIC_MAX_PEERS];        // acc_cur of every rank (own entry = local pointer)
  unsigned* peer_flags[JIC_MAX_PEERS];     // flag array of every rank; entry [r] is written by rank r
  unsigned long long* seq;                 // fused steps completed so far (local; identical on all ranks)
  int* error;                              // sticky: 3 = a peer did not arrive within the spin limit
};

constexpr long long kPeerSpinLimit = 16000000000ll;  // clock64 ticks (~8 s): a missing peer must not hang the GPU; the error is
                                                     // sticky, so the steps still queued behind a failure do not wait again

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// one node of a peer's raw grid (kAccRow = 4 reals, 32- / 16-byte aligned): vector loads, issued back to back by the caller for
// all peers before the first use so that the NVLink round trips overlap instead of adding up
__device__ __forceinline__ void ld_peer4(const double* p, double v[4]) {
  asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "l"(p) : "memory");
  asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(v[2]), "=d"(v[3]) : "l"(p + 2) : "memory");
}
__device__ __forceinline__ void ld_peer4(const float* p, float v[4]) {
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p) : "memory");
}

constexpr int kFieldsMcThreads = 512;

template <typename R>
__global__ void __launch_bounds__(kFieldsMcThreads) k_fields_mc(const FieldArgsMC<R> a) {
  extern __shared__ __align__(16) double fsm[];
  const int G = a.G, S = a.S, H = a.H, tid = threadIdx.x, nt = blockDim.x;
  const int lo = blockIdx.x * S, hi = min(lo + S, G);
  const int W = S + 2 * H;           // filter window: local li <-> global node lo - H + li
  const int Wm = S + 4;              // Maxwell window: local mi <-> global node lo - 2 + mi  (li = mi + H - 2)
  double* cur = fsm;                 // [4][W]
  double* nxt = fsm + 4 * W;         // [4][W]
  double* Em = fsm + 8 * W;          // [3][Wm]
  double* Bm = Em + 3 * Wm;          // [3][Wm]
  const bool periodic = (a.fbl == JIC_BC_PERIODIC) && (a.fbr == JIC_BC_PERIODIC);
  const int idx0 = H - lo;           // local (filter window) index of node 0
  const int idxL = G - 1 - lo + H;   // local (filter window) index of node G - 1
  const bool edgeL = !periodic && idx0 >= 0, edgeR = !periodic && idxL <= W - 1;  // the window reaches the domain edge
  const long long row = a.ctl->hist_row;

  // ---- N GPUs, fused: tell every rank that this rank's push is complete (it precedes this kernel in stream order), then
  //      wait for the same word from every rank.  One flag per (writer, reader) pair holding the step number, so nothing is
  //      ever reset.  Passing this point also proves that every rank has finished its PREVIOUS field kernel, i.e. nobody
  //      still reads the buffer this kernel is about to zero (acc_next).
  const bool fused = a.world > 1;
  if (fused) {
    const unsigned want = (unsigned)(*a.seq) + 1u;
    if (blockIdx.x == 0 && tid < a.world) st_release_sys(a.peer_flags[tid] + a.rank, want);
    if (tid < a.world && *(volatile int*)a.error != 3) {
      const unsigned* f = a.peer_flags[a.rank] + tid;
      const long long t0 = clock64();
      while ((int)(ld_acquire_sys(f) - want) < 0) {
        if (clock64() - t0 > kPeerSpinLimit) { atomicExch(a.error, 3); break; }
      }
    }
    __syncthreads();
  }
  // ---- raw grid window -> shared memory (periodic: wrapped; outside a non-periodic domain: unused)
  for (int li = tid; li < W; li += nt) {
    int g = lo - H + li;
    if (periodic) g = g < 0 ? g + G : (g >= G ? g - G : g);
    const bool in = g >= 0 && g < G;
    if (!fused) {
#pragma unroll
      for (int c = 0; c < kAccRow; ++c) cur[c * W + li] = in ? (double)a.acc_cur[g * kAccRow + c] : 0.0;
    } else {
      double sum[kAccRow] = {0.0, 0.0, 0.0, 0.0};
      if (in) {
        R vals[JIC_MAX_PEERS][kAccRow];
#pragma unroll
        for (int r = 0; r < JIC_MAX_PEERS; ++r)  // all loads first (own rank included: same path, local memory) ...
          if (r < a.world) ld_peer4(a.peer_acc[r] + g * kAccRow, vals[r]);
#pragma unroll
        for (int r = 0; r < JIC_MAX_PEERS; ++r)  // ... then the sum in rank order: identical rounding on every rank
          if (r < a.world) {
#pragma unroll
            for (int c = 0; c < kAccRow; ++c) sum[c] += (double)vals[r][c];
          }
      }
#pragma unroll
      for (int c = 0; c < kAccRow; ++c) cur[c * W + li] = sum[c];
    }
  }
  // the other raw buffer was consumed by the previous step: zero our slice of it for the next push
  for (int k = tid; k < (hi - lo) * kAccRow; k += nt) a.acc_next[lo * kAccRow + k] = R(0);
  if (a.ExC) for (int k = tid; k < hi - lo; k += nt) a.acc_next[(size_t)G * kAccRow + lo + k] = R(0);  // face component
  __syncthreads();

  // ---- digital filter on a shrinking valid range [va, vb); the domain edge (non-periodic) does not shrink
  int va = edgeL ? idx0 : 0, vb = edgeR ? idxL + 1 : W;
  if (a.passes > 0) {
    const int p_cl = a.passes < 17 ? a.passes : 17;
    const int n_reg = (a.passes - 1) < 16 ? (a.passes - 1) : 16;
    const double comp_alpha = p_cl - a.alpha * (p_cl - 1);
    const double zl = a.fbl == JIC_BC_ABSORBING ? 0.0 : 1.0, zr = a.fbr == JIC_BC_ABSORBING ? 0.0 : 1.0;
    for (int si = 0; si < a.n_strides; ++si) {
      const int s = a.strides[si];
      for (int sweep = 0; sweep <= n_reg; ++sweep) {
        const double al = sweep < n_reg ? a.alpha : comp_alpha;
        const double co = (1 - al) * 0.5;
        const int b0 = edgeL ? va : va + s, b1 = edgeR ? vb : vb - s;
        const int c = tid & 3;  // four components side by side: thread -> (component, every (nt/4)-th node)
        for (int li = b0 + (tid >> 2); li < b1; li += nt >> 2) {
          const double* y = cur + c * W;
          const double l = (edgeL && li - s < idx0) ? zl * y[idx0] : y[li - s];
          const double r = (edgeR && li + s > idxL) ? zr * y[idxL] : y[li + s];
          nxt[c * W + li] = al * y[li] + co * (l + r);
        }
        __syncthreads();
        double* t = cur; cur = nxt; nxt = t;
        va = b0; vb = b1;
      }
    }
  }
  // filtered J, rho of the owned nodes
  for (int i = tid; i < hi - lo; i += nt) {
    const int li = i + H, g = lo + i;
    a.J[g * 3 + 0] = cur[li]; a.J[g * 3 + 1] = cur[W + li]; a.J[g * 3 + 2] = cur[2 * W + li];
    a.rho[g] = cur[3 * W + li];
  }

  // ---- Maxwell window
  for (int mi = tid; mi < Wm; mi += nt) {
    int g = lo - 2 + mi;
    if (periodic) g = g < 0 ? g + G : (g >= G ? g - G : g);
    const bool in = g >= 0 && g < G;
#pragma unroll
    for (int c = 0; c < 3; ++c) { Em[c * Wm + mi] = in ? a.E_r[g * 3 + c] : 0.0; Bm[c * Wm + mi] = in ? a.B_r[g * 3 + c] : 0.0; }
  }
  __syncthreads();
  const double h = a.dt / 2;
  // window nodes that exist: [m_first, m_last]; wallL / wallR: the window starts / ends at a non-periodic wall
  const int m_first = periodic ? 0 : max(0, 2 - lo), m_last = periodic ? Wm - 1 : min(Wm - 1, G - 1 - lo + 2);
  const bool wallL = !periodic && lo - 2 + m_first == 0, wallR = !periodic && lo - 2 + m_last == G - 1;
  auto faraday_mc = [&](int from, int to) {  // B -= h curl E on mi in [from, to]   (_fields.py:102-111)
    for (int mi = from + tid; mi <= to; mi += nt) {
      double ey_m, ez_m;
      if (wallL && mi == m_first) {  // left ghost of E (_boundary_conditions.py:148-176)
        if (a.fbl == JIC_BC_REFLECTIVE) { ey_m = Em[Wm + mi]; ez_m = Em[2 * Wm + mi]; }
        else { ey_m = -2 * kC * Bm[2 * Wm + mi] - Em[Wm + mi]; ez_m = 2 * kC * Bm[Wm + mi] - Em[2 * Wm + mi]; }
      } else { ey_m = Em[Wm + mi - 1]; ez_m = Em[2 * Wm + mi - 1]; }
      const double dFz = (Em[2 * Wm + mi] - ez_m) / a.dx, dFy = (Em[Wm + mi] - ey_m) / a.dx;
      Bm[Wm + mi] -= h * (-dFz);
      Bm[2 * Wm + mi] -= h * dFy;
    }
    __syncthreads();
  };
  auto ampere_mc = [&](int from, int to) {  // E += h (c^2 curl B - J/eps0) on mi in [from, to]   (_fields.py:132-144)
    for (int mi = from + tid; mi <= to; mi += nt) {
      const int li = mi + H - 2;
      double by_p, bz_p;
      if (wallR && mi == m_last) {  // right ghost of B (_boundary_conditions.py:178-207)
        if (a.fbr == JIC_BC_REFLECTIVE) { by_p = Bm[Wm + mi]; bz_p = Bm[2 * Wm + mi]; }
        else { by_p = -(2 / kC) * Em[2 * Wm + mi] - Bm[Wm + mi]; bz_p = (2 / kC) * Em[Wm + mi] - Bm[2 * Wm + mi]; }
      } else { by_p = Bm[Wm + mi + 1]; bz_p = Bm[2 * Wm + mi + 1]; }
      const double dFz = (bz_p - Bm[2 * Wm + mi]) / a.dx, dFy = (by_p - Bm[Wm + mi]) / a.dx;
      Em[mi] += h * ((kC * kC) * 0.0 - (cur[li] / kEps0));
      Em[Wm + mi] += h * ((kC * kC) * (-dFz) - (cur[W + li] / kEps0));
      Em[2 * Wm + mi] += h * ((kC * kC) * dFy - (cur[2 * W + li] / kEps0));
    }
    __syncthreads();
  };
  // second half step of step n: B then E (_fields.py:185-193).  Valid ranges shrink by one node per neighbour access.
  const int fa = wallL ? m_first : m_first + 1;     // B valid on [fa, m_last]
  faraday_mc(fa, m_last);
  const int ab = wallR ? m_last : m_last - 1;       // E valid on [fa, ab]
  ampere_mc(fa, ab);
  if (a.ExC) {  // _algorithms.py:78: E_x is replaced by the Gauss / Poisson solve
    for (int mi = fa + tid; mi <= ab; mi += nt) {
      int g = lo - 2 + mi;
      if (periodic) g = g < 0 ? g + G : (g >= G ? g - G : g);
      Em[mi] = a.ExC[g];
    }
    __syncthreads();
  }
  {  // step outputs (_algorithms.py:93) for the owned nodes
    R* hE = a.record ? (R*)a.ctl->hist[0] : nullptr;
    R* hB = a.record ? (R*)a.ctl->hist[1] : nullptr;
    R* hJ = a.record ? (R*)a.ctl->hist[2] : nullptr;
    R* hrho = a.record ? (R*)a.ctl->hist[3] : nullptr;
    for (int k = tid; k < (hi - lo) * 3; k += nt) {
      const int i = k / 3, c = k - 3 * i, mi = i + 2, g = lo + i;
      const double e = Em[c * Wm + mi], b = Bm[c * Wm + mi];
      a.E_int[g * 3 + c] = e; a.B_int[g * 3 + c] = b;
      if (hE) hE[(size_t)row * G * 3 + g * 3 + c] = (R)e;
      if (hB) hB[(size_t)row * G * 3 + g * 3 + c] = (R)b;
      if (hJ) hJ[(size_t)row * G * 3 + g * 3 + c] = (R)cur[c * W + i + H];
    }
    if (hrho) for (int i = tid; i < hi - lo; i += nt) hrho[(size_t)row * G + lo + i] = (R)cur[3 * W + i + H];
  }
  // first half step of step n+1: E then B (_fields.py:175-183)
  ampere_mc(fa, ab);
  const int fb = wallL ? fa : fa + 1;               // B valid on [fb, ab]
  faraday_mc(fb, ab);
  // ---- new leap-frog state and the padded total-field table of the owned nodes (+ the ghost rows they source)
  for (int i = tid; i < hi - lo; i += nt) {
    const int mi = i + 2, g = lo + i;
    double tE[3], tB[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double e = Em[c * Wm + mi], b = Bm[c * Wm + mi];
      a.E_w[g * 3 + c] = e; a.B_w[g * 3 + c] = b;
      tE[c] = e + a.extE[g * 3 + c]; tB[c] = b + a.extB[g * 3 + c];
    }
    auto put = [&](int r) {
      R* f = a.F + (size_t)r * kFieldRow;
#pragma unroll
      for (int c = 0; c < 3; ++c) { f[c] = (R)tE[c]; f[3 + c] = (R)tB[c]; }
      f[6] = R(0); f[7] = R(0);
    };
    put(g + 2);
    // rows [L2, L1 | ... | R]: periodic f[G-2], f[G-1] | f[0]; reflective f[1], f[0] | f[G-1]; absorbing zeros (written below)
    if (a.fbl == JIC_BC_PERIODIC) { if (g == G - 2) put(0); if (g == G - 1) put(1); }
    else if (a.fbl == JIC_BC_REFLECTIVE) { if (g == 1) put(0); if (g == 0) put(1); }
    if (a.fbr == JIC_BC_PERIODIC) { if (g == 0) put(G + 2); }
    else if (a.fbr == JIC_BC_REFLECTIVE) { if (g == G - 1) put(G + 2); }
  }
  if (blockIdx.x == 0 && tid < 3 * kFieldRow) {  // absorbing ghost rows are zeros
    const int r = tid / kFieldRow, c = tid - r * kFieldRow;
    if (r < 2 && a.fbl == JIC_BC_ABSORBING) a.F[(size_t)r * kFieldRow + c] = R(0);
    if (r == 2 && a.fbr == JIC_BC_ABSORBING) a.F[(size_t)(G + 2) * kFieldRow + c] = R(0);
  }
  // ---- the last CTA to finish advances the run counters (every CTA has read hist_row by then)
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const unsigned k = atomicAdd(a.done, 1u);
    if (k == gridDim.x - 1) {
      a.ctl->hist_row = row + 1;
      a.ctl->step += 1;
      if (fused) *a.seq += 1ull;
      *a.done = 0u;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// K2g  per-step electrostatic correction (field_solver != 0, jaxincell/_algorithms.py:69-78): E_x from the filtered charge
//      density on the faces.  The two spectral solvers (_fields.py:9-60) are the same circulant operator, E = h (*) rho with
//      h[d] = 1/(G eps0) sum_{m != 0} sin(2 pi m d / G) / k_m  (k = 0 and Nyquist terms have no real part), evaluated here as a
//      direct circular convolution: G^2 multiply-adds spread over the SMs is ~1 us at G = 4096 and needs no transform library
//      inside the captured graph.  field_solver = 2 (_fields.py:62-81) is the prefix sum (dx/eps0) sum_{j<=i} rho_j.
//      Every CTA filters the whole face component itself (O(15 G), shared memory) and then produces its own slice of E_x,
//      one warp per node, lanes striding over the sources.
// ---------------------------------------------------------------------------------------------------------
__global__ void k_gauss_kernel(int G, double dx, double* h) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= G) return;
  const int n_pos = (G - 1) / 2 + 1;  // numpy/jnp fftfreq: indices [0, n_pos) are >= 0, the rest are m - G
  double sum = 0.0;
  for (int m = 1; m < G; ++m) {
    const int mm = m < n_pos ? m : m - G;
    const double k = ((double)mm * (1.0 / ((double)G * dx))) * 2.0 * 3.14159265358979323846;
    const long long ph = ((long long)m * d) % G;
    sum += sinpi(2.0 * (double)ph / (double)G) / k;
  }
  h[d] = sum / ((double)G * kEps0);
}

template <typename R>
struct GaussArgs {
  int G, fbl, fbr, passes, n_strides, mode;  // mode = field_solver (1 Gauss FFT, 2 Gauss Cartesian, 3 Poisson FFT)
  int strides[JIC_MAX_STRIDES];
  double alpha, dx;
  const R* accS;    // raw rho on the faces (already all-reduced)
  const double* h;  // circulant kernel of the spectral solvers (G)
  double* Ex;       // out: E_x (G)
};

constexpr int kGaussThreads = 256;

template <typename R>
__global__ void __launch_bounds__(kGaussThreads) k_gauss(const GaussArgs<R> a) {
  extern __shared__ __align__(16) double gsm[];  // 2 G doubles
  const int G = a.G, tid = threadIdx.x, nt = blockDim.x;
  for (int j = tid; j < G; j += nt) gsm[j] = (double)a.accS[j];
  __syncthreads();
  const double* rho = filter_components<double*>(gsm, gsm + G, 1, G, a.passes, a.alpha, a.n_strides, a.strides, a.fbl, a.fbr);
  const int S = (G + gridDim.x - 1) / gridDim.x, lo = blockIdx.x * S, hi = min(lo + S, G);
  const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  for (int i = lo + warp; i < hi; i += nw) {
    double sum = 0.0;
    if (a.mode == 2) {
      for (int j = lane; j <= i; j += 32) sum += rho[j];
      sum *= a.dx / kEps0;
    } else {
      // h[(i - j) mod G]: lanes read consecutive (descending) addresses
      for (int j = lane; j < G; j += 32) {
        int k = i - j;
        k += k < 0 ? G : 0;
        sum = fma(rho[j], __ldg(a.h + k), sum);
      }
    }
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) a.Ex[i] = sum;
  }
}

// ---------------------------------------------------------------------------------------------------------
// small utilities
// ---------------------------------------------------------------------------------------------------------
template <typename R>
__global__ void k_export_particles(const DevParams<R> p, const R* xh, const R* yh, const R* zh, const R* vx, const R* vy, const R* vz,
                                   R* x_out, R* v_out, uint8_t* alive) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.N; i += (long long)gridDim.x * blockDim.x) {
    const R x = xh[i];
    if (x_out) { x_out[3 * i] = x; x_out[3 * i + 1] = p.track_yz ? yh[i] : R(0); x_out[3 * i + 2] = p.track_yz ? zh[i] : R(0); }
    if (v_out) { v_out[3 * i] = vx[i]; v_out[3 * i + 1] = vy[i]; v_out[3 * i + 2] = vz[i]; }
    if (alive) alive[i] = !((x < -p.half_L) || (x > p.half_L));
  }
}

template <typename R>
__global__ void k_kinetic(const DevParams<R> p, const R* vx, const R* vy, const R* vz, double* out) {
  double acc = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.N; i += (long long)gridDim.x * blockDim.x) {
    const int s = species_of(i, p);
    const double a = vx[i], b = vy[i], c = vz[i];
    acc += 0.5 * (double)p.sp_m[s] * (a * a + b * b + c * c);
  }
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

// per-species kinetic energy of the step into row hist_row of the (T, n_species) history (jic_outputs.kinetic_energy), INDEXED / CN
// layouts: the species of particle i follows from its index.  Enqueued between the push and the field kernel of a step.
template <typename R>
__global__ void __launch_bounds__(256) k_kinetic_hist(const DevParams<R> p, const R* vx, const R* vy, const R* vz, const RunControl* ctl, int row_back) {
  double* out = (double*)ctl->hist[6];
  if (!out) return;
  out += (ctl->hist_row - row_back) * p.n_species;  // (row_back = 1: the Crank-Nicolson field kernel has already advanced the row)
  double acc[JIC_MAX_SPECIES];
#pragma unroll
  for (int s = 0; s < JIC_MAX_SPECIES; ++s) acc[s] = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.N; i += (long long)gridDim.x * blockDim.x) {
    const int s = species_of(i, p);
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
      if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(out + s, v);
    }
  }
}

// start of a jic_run: row 0 of the caller's history buffers is the first step of this call
__global__ void k_begin_run(RunControl* ctl, jic_outputs out) {
  ctl->hist_row = 0;
  ctl->hist[0] = out.electric_field; ctl->hist[1] = out.magnetic_field; ctl->hist[2] = out.current_density;
  ctl->hist[3] = out.charge_density; ctl->hist[4] = out.positions; ctl->hist[5] = out.velocities; ctl->hist[6] = out.kinetic_energy;
}

template <typename R>
__global__ void k_f32_to_f64(const float* src, double* dst, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src ? (double)src[i] : 0.0;
}

template <typename Src, typename Dst>
__global__ void k_convert(const Src* src, Dst* dst, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = (Dst)src[i];
}

}  // namespace jic
