// Loading the reference's scan carry into the INDEXED engine (jic_load_carry): the step-granularity boundary of SURVEY.md 8(b).
//
// The reference's carry (jaxincell/_simulation.py:228-231, consumed by Boris_step at _algorithms.py:23-24) is
//   (E^n, B^n, x_{n-1/2}, x_n, x_{n+1/2}, v_n, q, m, q/m);
// the engine carries x_{n+1/2}, v_n, the FILTERED current of the last deposit, the fields advanced by the first Maxwell half step of
// the coming step, and the padded gather table.  Two kernels rebuild that state; they only call device functions that the verified
// kernels already use (make_cloud, deposit_jx, deposit_cloud, ampere, faraday) and leave those kernels untouched:
//   k_load_carry    particles -> SoA state, raw J of (x_{n-1/2}, x_n, x_{n+1/2}, v_n)   (the deposit of _algorithms.py:29-32)
//   [k_fields in init mode filters the raw grid -> J; its Gauss solve / zero B are overwritten next]
//   k_carry_fields  E^n, B^n in, then E then B by dt/2 (_fields.py:175-183) and the gather table (_algorithms.py:36-43)
#pragma once
#include "jic_kernels.cuh"
#include "jic_cn.cuh"

namespace jic {

template <typename R>
__global__ void __launch_bounds__(256) k_load_carry(const DevParams<R> p, const R* __restrict__ x_minus, const R* __restrict__ x_n,
                                                    const R* __restrict__ x_plus, const R* __restrict__ v_n, R* __restrict__ xh,
                                                    R* __restrict__ yh, R* __restrict__ zh, R* __restrict__ vx, R* __restrict__ vy,
                                                    R* __restrict__ vz, R* __restrict__ v_init, R* __restrict__ acc) {
  const GlobalGrid<R> grid{acc};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.N; i += (long long)gridDim.x * blockDim.x) {
    const R xp = x_plus[3 * i], xm = x_minus[3 * i], x0 = x_n[3 * i];
    const R v[3] = {v_n[3 * i], v_n[3 * i + 1], v_n[3 * i + 2]};
    // absorbed particles sit parked outside the box with q = 0 (_boundary_conditions.py:40,51,64,78): k_step recognises them the same way
    const bool dead = (xp < -p.half_L) || (xp > p.half_L);
    if (!dead) {
      const int s = species_of(i, p);
      const R q = p.sp_q[s], a = q * p.inv_dx;
      const Cloud<R> cm = make_cloud(xm, p), cp = make_cloud(xp, p), c0 = make_cloud(x0, p);
      deposit_jx_startup(grid, xm, cm, cp, q / p.dt, p);
      deposit_cloud(grid, c0, p.G, a * v[1], a * v[2], a, true);  // J_y,z = rho(x_n) v_{y,z}; the rho component is not used afterwards
    }
    xh[i] = xp;
    vx[i] = v[0]; vy[i] = v[1]; vz[i] = v[2];
    if (p.track_yz) { yh[i] = x_plus[3 * i + 1]; zh[i] = x_plus[3 * i + 2]; }
    if (v_init) { v_init[3 * i] = v[0]; v_init[3 * i + 1] = v[1]; v_init[3 * i + 2] = v[2]; }
  }
}

template <typename R>
__global__ void __launch_bounds__(1024) k_carry_fields(const FieldArgs<R> a, const R* __restrict__ E_in, const R* __restrict__ B_in) {
  const int G = a.G, tid = threadIdx.x, nt = blockDim.x;
  for (int k = tid; k < G * 3; k += nt) {
    const double e = (double)E_in[k], b = (double)B_in[k];
    a.E[k] = e; a.B[k] = b; a.E_int[k] = e; a.B_int[k] = b; a.E0[k] = e; a.B0[k] = b;
  }
  __syncthreads();
  const double h = a.dt / 2;
  // first half step of the coming step: E then B (_fields.py:175-183), with the filtered J that k_fields left in a.J
  ampere(a.E, a.B, a.J, G, a.fbr, a.dx, h);
  faraday(a.E, a.B, G, a.fbl, a.dx, h);
  // padded total fields for the gather: rows [L2, L1, f_0..f_{G-1}, R]  (same table as step 5 of k_fields)
  for (int r = tid; r < G + 3; r += nt) {
    int src;
    if (r >= 2 && r < G + 2) src = r - 2;
    else if (r < 2) src = a.fbl == JIC_BC_PERIODIC ? (G - 2 + r) : a.fbl == JIC_BC_REFLECTIVE ? (1 - r) : -1;
    else src = a.fbr == JIC_BC_PERIODIC ? 0 : a.fbr == JIC_BC_REFLECTIVE ? (G - 1) : -1;
    if (src >= G) src = G - 1;
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

// ---- field_solver != 0, step 0 only.  The step kernels deposit rho(x_n) on the faces from x_n = BCpos(x_{n+1/2} - dt/2 v_n), which is what the
// reference holds in `positions` from step 1 on (_algorithms.py:60-61 of the previous step) -- but in step 0 `positions` is the raw x_0
// (_simulation.py:228-231).  The two agree unless the start-up half step took the particle through a PERIODIC wall while the opposite wall
// is not periodic (the way back is then reflected or absorbed instead of wrapped).  For exactly those particles this kernel, launched
// once after the start-up kernel, puts q [S(x_0) - S(x_n)] on the faces, so that step 0 ends up with the reference's deposit.  Found by
// running the CUDA source on the CPU against the reference semantics over random boundary combinations (tests/test_cuda_source_on_cpu.py).
// (A particle that also gets absorbed during step 0 keeps this correction although the reference drops its charge: second order, ignored.)
template <typename R>
__global__ void __launch_bounds__(256) k_start_face_fix(const DevParams<R> p, const R* __restrict__ x0, const R* __restrict__ v0, long long i0, long long n,
                                                        R* __restrict__ acc) {
  const GlobalGrid<R> grid{acc, p.G};
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
    const R X0 = x0[3 * j];
    R vx0 = v0[3 * j];
    R xp = X0 + p.half_dt * vx0;
    const int flag = bc_x(xp, p);
    if (flag == 2) continue;  // absorbed by the start-up half step: no charge
    if (flag == 1) vx0 = -vx0;
    R xn = xp - p.half_dt * vx0;  // what k_step / k_push take for x_n in step 0
    bc_x(xn, p);
    if (fabs(xn - X0) > R(1e-6) * p.dx) {
      const R a = p.sp_q[species_of(i0 + j, p)] * p.inv_dx;
      deposit_faces(grid, make_cloud_faces(X0, p), p.G, a);
      deposit_faces(grid, make_cloud_faces(xn, p), p.G, -a);
    }
  }
}

// ---- Crank-Nicolson: the carry of CN_step is (E, B, x_n, v_n, q, m, q/m) (jaxincell/_simulation.py:237-240, _algorithms.py:103-104).
// Particles and fields are copied as they are; `alive_in` (0 = the particle's q is zero in the carry: absorbed by the start-up half
// step, _simulation.py:217-220) replaces the byte k_cn_start computes.  The averaged tables of the first Picard iteration are then
// prepared by k_cn_fields exactly as after jic_initialize.
template <typename R>
__global__ void __launch_bounds__(256) k_cn_load(const DevParams<R> p, const R* __restrict__ x, const R* __restrict__ v, const uint8_t* __restrict__ alive_in,
                                                 CnState<R> s, R* __restrict__ v_init, uint8_t* __restrict__ alive) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.N; i += (long long)gridDim.x * blockDim.x) {
    s.x[i] = x[3 * i]; s.y[i] = x[3 * i + 1]; s.z[i] = x[3 * i + 2];
    s.vx[i] = v[3 * i]; s.vy[i] = v[3 * i + 1]; s.vz[i] = v[3 * i + 2];
    alive[i] = alive_in ? (alive_in[i] != 0) : 1;
    if (v_init) { v_init[3 * i] = v[3 * i]; v_init[3 * i + 1] = v[3 * i + 1]; v_init[3 * i + 2] = v[3 * i + 2]; }
  }
}

template <typename R>
__global__ void k_carry_copy_fields(const R* __restrict__ E_in, const R* __restrict__ B_in, double* E, double* B, double* E0, double* B0, int n) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const double e = (double)E_in[k], b = (double)B_in[k];
    E[k] = e; B[k] = b; E0[k] = e; B0[k] = b;
  }
}

}  // namespace jic
