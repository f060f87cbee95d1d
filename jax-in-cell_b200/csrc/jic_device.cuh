// Device-side building blocks of the JAX-in-Cell Boris hot path (sm_100a).
//
// Every function states the reference code whose arithmetic it reproduces (paths relative to the reference
// checkout).  They are written per particle on the particle's own 3-node / 6-node stencil: the reference's
// O(N*G) "each particle builds a length-G vector" formulation (jaxincell/_sources.py:101-104, :199-207) is an
// artefact of expressing the scheme in array primitives, not part of the scheme.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>

#include "../../include/jic_b200.h"

namespace jic {

constexpr double kEps0 = 8.85418782e-12;   // jaxincell/_constants.py:1
constexpr double kMu0 = 1.25663706e-6;     // :2
constexpr double kC = 2.99792458e8;        // :3

constexpr int JIC_MAX_PEERS = 8;  // GPUs of one NVSwitch box

// Row stride (in reals) of the padded total-field table the gather reads: Ex,Ey,Ez,Bx,By,Bz,pad,pad.
constexpr int kFieldRow = 8;
// Row stride of the raw deposition grid: Jx,Jy,Jz,rho.
constexpr int kAccRow = 4;

struct RunControl {
  long long step;      // steps completed since jic_initialize
  long long hist_row;  // row of the history buffers the next step writes
  void* hist[7];       // jic_outputs of the current jic_run (E, B, J, rho, positions, velocities, kinetic energy); read by the kernels at
                       // run time so that one captured graph serves every set of output buffers
  // device-side timing of the binned push kernel (%globaltimer): first CTA in, last CTA out of the current launch; the kernel behind
  // it folds the difference into the running sum (jic_push_kernel_time) -- valid inside CUDA-graph replays, where events are not
  unsigned long long push_t0, push_t1, push_ns, push_launches;
};

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <typename R>
struct DevParams {
  long long N;
  int G;
  int n_species;
  int pbl, pbr, fbl, fbr;
  int relativistic;
  int track_yz;
  int stag;        // field_solver != 0: also deposit rho(x_n) on the faces g_k + dx/2 (_algorithms.py:69-72)
  int park_left_cell;  // cell of park_left as the reference's float `//` gives it (reference_floor_div): the start-up J_x window of a
                       // particle whose x_{-1/2} was parked while it keeps its charge (_simulation.py:216-225)
  R L, Ly, Lz, half_L, half_Ly, half_Lz;
  R dx, inv_dx, half_dx, dt, half_dt;
  R g0, gl;        // grid[0], grid[-1]
  R gs;            // grid[0] - dx/2  (grid_start of the current deposit, _algorithms.py:30)
  R park_left, park_right;  // where absorbed particles are parked (_boundary_conditions.py:40,51)
  long long sp_end[JIC_MAX_SPECIES];
  R sp_q[JIC_MAX_SPECIES], sp_m[JIC_MAX_SPECIES], sp_qm[JIC_MAX_SPECIES];
};

// jnp.floor_divide on floats -- the reference's `//` in cell_no = (x - grid_start) // dx (_sources.py:190): the floor of the exact quotient
// of the two floats, through the remainder (host side; used for DevParams::park_left_cell)
inline int reference_floor_div(double a, double b) {
  const double mod = std::fmod(a, b);
  double div = (a - mod) / b;
  if (mod != 0.0 && ((b < 0.0) != (mod < 0.0))) div -= 1.0;
  return (int)std::nearbyint(div);
}

template <typename R>
__device__ __forceinline__ R floor_mod(R a, R b) {  // XLA / NumPy float `%` for b > 0
  R r = fmod(a, b);
  if (r < R(0)) r += b;
  return r;
}

// ---------------------------------------------------------------------------------------------------------
// Particle boundaries: jaxincell/_boundary_conditions.py:32-56 (x), :28-29 (y,z).  Returns 0 = untouched or
// periodic wrap, 1 = reflected (v_x flips), 2 = absorbed (v = 0, q = q/m = 0).  Strict inequalities.
// ---------------------------------------------------------------------------------------------------------
template <typename R>
__device__ __forceinline__ int bc_x(R& x, const DevParams<R>& p) {
  if (x < -p.half_L) {
    if (p.pbl == JIC_BC_PERIODIC) { x = floor_mod(x + p.half_L, p.L) - p.half_L; return 0; }
    if (p.pbl == JIC_BC_REFLECTIVE) { x = -p.L - x; return 1; }
    x = p.park_left; return 2;
  }
  if (x > p.half_L) {
    if (p.pbr == JIC_BC_PERIODIC) { x = floor_mod(x + p.half_L, p.L) - p.half_L; return 0; }
    if (p.pbr == JIC_BC_REFLECTIVE) { x = p.L - x; return 1; }
    x = p.park_right; return 2;
  }
  return 0;
}

template <typename R>
__device__ __forceinline__ R wrap_transverse(R y, R len, R half_len) {
  const R t = y + half_len;
  if (t >= R(0) && t < len) return t - half_len;  // (fmod returns its first argument bit for bit on [0, len): skip the slow remainder loop)
  return floor_mod(t, len) - half_len;
}

// ---------------------------------------------------------------------------------------------------------
// Quadratic-spline (S2) cloud of one particle on cell centres: jaxincell/_sources.py:83-110 with the ghost
// fold of charge_density_BCs (:43-81).  Weights are dimensionless (the q/dx factor is applied by the caller).
//   nodes c-1, c, c+1 get w[0..2] when they are on the grid; `first` / `last` is what nodes 0 / G-1 receive
//   additionally from the folded ghost weight.
// ---------------------------------------------------------------------------------------------------------
template <typename R>
struct Cloud {
  int c;
  R w[3];
  R first, last;
};

template <typename R>
__device__ __forceinline__ Cloud<R> make_cloud(R x, const DevParams<R>& p) {
  Cloud<R> cl;
  const R s = (x - p.g0) * p.inv_dx;
  int c = (int)floor(s + R(0.5));
  const bool inside = (x >= -p.half_L) && (x <= p.half_L);
  if (inside) c = min(max(c, 0), p.G - 1);
  const R d = s - R(c);
  cl.c = c;
  cl.w[0] = R(0.5) * (R(0.5) - d) * (R(0.5) - d);
  cl.w[1] = R(0.75) - d * d;
  cl.w[2] = R(0.5) * (R(0.5) + d) * (R(0.5) + d);
  // the ghost weight is switched by the reference's own test |x - g_0| <= dx/2 (resp. g_{G-1})
  R exl = R(0), exr = R(0);
  if (fabs(x - p.g0) <= p.half_dx) { const R t = R(0.5) + (p.g0 - x) * p.inv_dx; exl = R(0.5) * t * t; }
  if (fabs(x - p.gl) <= p.half_dx) { const R t = R(0.5) + (x - p.gl) * p.inv_dx; exr = R(0.5) * t * t; }
  cl.first = (p.pbl == JIC_BC_PERIODIC ? exr : R(0)) + (p.pbl == JIC_BC_REFLECTIVE ? exl : R(0));
  cl.last = (p.pbr == JIC_BC_PERIODIC ? exl : R(0)) + (p.pbr == JIC_BC_REFLECTIVE ? exr : R(0));
  return cl;
}

template <typename R>
__device__ __forceinline__ R cloud_at(const Cloud<R>& cl, int k, int G) {
  const int j = k - cl.c;
  R v = (j == -1) ? cl.w[0] : (j == 0) ? cl.w[1] : (j == 1) ? cl.w[2] : R(0);
  if (k == 0) v += cl.first;
  if (k == G - 1) v += cl.last;
  return v;
}

// The same cloud on the FACES g_k + dx/2 (`grid + dx/2` of jaxincell/_algorithms.py:70).  The nearest face of a particle in the
// left half cell is face -1: no clamp, its weights are dropped like the reference's `where` over the grid does, and the ghost
// fold only exists for |x - f_0| <= dx/2 resp. |x - f_{G-1}| <= dx/2 (_sources.py:60-69 with the shifted grid).
template <typename R>
__device__ __forceinline__ Cloud<R> make_cloud_faces(R x, const DevParams<R>& p) {
  Cloud<R> cl;
  const R f0 = p.g0 + p.half_dx, fl = p.gl + p.half_dx;
  const R s = (x - f0) * p.inv_dx;
  const int c = (int)floor(s + R(0.5));
  const R d = s - R(c);
  cl.c = c;
  cl.w[0] = R(0.5) * (R(0.5) - d) * (R(0.5) - d);
  cl.w[1] = R(0.75) - d * d;
  cl.w[2] = R(0.5) * (R(0.5) + d) * (R(0.5) + d);
  R exl = R(0), exr = R(0);
  if (fabs(x - f0) <= p.half_dx) { const R t = R(0.5) + (f0 - x) * p.inv_dx; exl = R(0.5) * t * t; }
  if (fabs(x - fl) <= p.half_dx) { const R t = R(0.5) + (x - fl) * p.inv_dx; exr = R(0.5) * t * t; }
  cl.first = (p.pbl == JIC_BC_PERIODIC ? exr : R(0)) + (p.pbl == JIC_BC_REFLECTIVE ? exl : R(0));
  cl.last = (p.pbr == JIC_BC_PERIODIC ? exl : R(0)) + (p.pbr == JIC_BC_REFLECTIVE ? exr : R(0));
  return cl;
}

__device__ __forceinline__ int mod_pos(int a, int n) {
  int r = a % n;
  return r < 0 ? r + n : r;
}

// Deposition targets -----------------------------------------------------------------------------------------
template <typename R>
struct GlobalGrid {  // straight to the L2-resident raw grid (RED.ADD.F64 / .F32)
  R* acc;
  int G = 0;  // only needed for add_face (the face component sits behind the (G,4) block)
  __device__ __forceinline__ void add(int node, int comp, R v) const { atomicAdd(acc + node * kAccRow + comp, v); }
  __device__ __forceinline__ void add_face(int node, R v) const { atomicAdd(acc + (size_t)G * kAccRow + node, v); }
};

template <typename R>
struct SharedGrid {  // CTA-private copy of the raw grid in shared memory, flushed once per CTA
  R* sacc;
  __device__ __forceinline__ void add(int node, int comp, R v) const { atomicAdd(sacc + node * kAccRow + comp, v); }
};

// rho-type deposit of (a_x?, a1, a2, a3) = per-component amplitudes on the nodes of one cloud.
template <typename R, typename Grid>
__device__ __forceinline__ void deposit_cloud(const Grid& g, const Cloud<R>& cl, int G, R ay, R az, R arho, bool with_j) {
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int k = cl.c + j - 1;
    if (k >= 0 && k < G) {
      if (with_j) { g.add(k, 1, cl.w[j] * ay); g.add(k, 2, cl.w[j] * az); }
      g.add(k, 3, cl.w[j] * arho);
    }
  }
  if (cl.first != R(0)) {
    if (with_j) { g.add(0, 1, cl.first * ay); g.add(0, 2, cl.first * az); }
    g.add(0, 3, cl.first * arho);
  }
  if (cl.last != R(0)) {
    if (with_j) { g.add(G - 1, 1, cl.last * ay); g.add(G - 1, 2, cl.last * az); }
    g.add(G - 1, 3, cl.last * arho);
  }
}

// rho(x_n) on the faces into the fifth raw component (field_solver != 0, jaxincell/_algorithms.py:69-72)
template <typename R, typename Grid>
__device__ __forceinline__ void deposit_faces(const Grid& g, const Cloud<R>& cl, int G, R a) {
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int k = cl.c + j - 1;
    if (k >= 0 && k < G) g.add_face(k, cl.w[j] * a);
  }
  if (cl.first != R(0)) g.add_face(0, cl.first * a);
  if (cl.last != R(0)) g.add_face(G - 1, cl.last * a);
}

// Charge-conserving J_x: jaxincell/_sources.py:190-207.  Window of min(6,G) nodes starting three nodes left of
// the cell of x_old, taken with a periodic roll whatever the BC; J_x[k_j] = -(q/dt) * sum_{i<=j}(w_new - w_old)[k_i].
// Whatever falls outside the window is dropped, as in the reference.
template <typename R, typename Grid>
__device__ __forceinline__ void deposit_jx_from_cell(const Grid& g, int cell, const Cloud<R>& c_old, const Cloud<R>& c_new, R q_over_dt,
                                                     const DevParams<R>& p) {
  const int W = p.G < 6 ? p.G : 6;
  R run = R(0);
  for (int j = 0; j < W; ++j) {
    const int k = mod_pos(cell - 3 + j, p.G);
    run += cloud_at(c_new, k, p.G) - cloud_at(c_old, k, p.G);
    if (run != R(0)) g.add(k, 0, -q_over_dt * run);
  }
}
template <typename R, typename Grid>
__device__ __forceinline__ void deposit_jx(const Grid& g, R x_old, const Cloud<R>& c_old, const Cloud<R>& c_new, R q_over_dt,
                                           const DevParams<R>& p) {
  deposit_jx_from_cell(g, (int)floor((x_old - p.gs) * p.inv_dx), c_old, c_new, q_over_dt, p);
}
// Start-up only (_simulation.py:216-225): x_{-1/2} = BCpos(x_0 - dt/2 v) can be PARKED at grid[0] - 1.5 dx while the particle keeps its
// charge (only the forward half step zeroes it).  That position sits exactly on a cell border of the J_x window, so the cell is whatever
// the reference's float floor-division makes of it -- computed once on the host with the same arithmetic (park_left_cell).  (The right
// park, grid[-1] + 3 dx, is half a cell away from a border; inside a run a parked x_{n-1/2} always comes with q = 0.)
template <typename R, typename Grid>
__device__ __forceinline__ void deposit_jx_startup(const Grid& g, R x_old, const Cloud<R>& c_old, const Cloud<R>& c_new, R q_over_dt,
                                                   const DevParams<R>& p) {
  const int cell = x_old == p.park_left ? p.park_left_cell : (int)floor((x_old - p.gs) * p.inv_dx);
  deposit_jx_from_cell(g, cell, c_old, c_new, q_over_dt, p);
}

// ---------------------------------------------------------------------------------------------------------
// Gather: jaxincell/_particles.py:29-45 as called from _algorithms.py:40-43.  `F` is the padded total field
// [L2, L1, f_0 .. f_{G-1}, R] x {Ex,Ey,Ez,Bx,By,Bz,-,-} built by the field kernel, so no BC logic is needed here.
//   E: faces g_k + dx/2, m = floor((x-g_0)/dx) in [-1, G-1];  B: centres, m = floor((x-g_0)/dx + 1/2).
// ---------------------------------------------------------------------------------------------------------
template <typename R>
__device__ __forceinline__ void gather_fields(const R* __restrict__ F, R x, const DevParams<R>& p, R E[3], R B[3]) {
  const R s = (x - p.g0) * p.inv_dx;
  const int rows = p.G + 3;
  {
    const R fl = floor(s);
    const R d = s - fl - R(0.5);
    int r0 = (int)fl + 1;  // padded row of f[m-1]
    r0 = min(max(r0, 0), rows - 3);
    const R w0 = R(0.5) * (R(0.5) - d) * (R(0.5) - d), w1 = R(0.75) - d * d, w2 = R(0.5) * (R(0.5) + d) * (R(0.5) + d);
    const R* f = F + (size_t)r0 * kFieldRow;
#pragma unroll
    for (int c = 0; c < 3; ++c) E[c] = w0 * __ldg(f + c) + w1 * __ldg(f + kFieldRow + c) + w2 * __ldg(f + 2 * kFieldRow + c);
  }
  {
    const R fl = floor(s + R(0.5));
    const R d = s - fl;
    int r0 = (int)fl + 1;
    r0 = min(max(r0, 0), rows - 3);
    const R w0 = R(0.5) * (R(0.5) - d) * (R(0.5) - d), w1 = R(0.75) - d * d, w2 = R(0.5) * (R(0.5) + d) * (R(0.5) + d);
    const R* f = F + (size_t)r0 * kFieldRow + 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) B[c] = w0 * __ldg(f + c) + w1 * __ldg(f + kFieldRow + c) + w2 * __ldg(f + 2 * kFieldRow + c);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Pushers: jaxincell/_particles.py:68-127 (Boris) and :132-200 (relativistic, momentum form).
// ---------------------------------------------------------------------------------------------------------
template <typename R>
__device__ __forceinline__ void boris_velocity(R v[3], const R E[3], const R B[3], R qm, R dt) {
  const R h = qm * dt * R(0.5);
  R vm[3] = {v[0] + h * E[0], v[1] + h * E[1], v[2] + h * E[2]};
  // R = v^- + (dt/2)(q/m) v^- x B ;  t = (q/m)(dt/2) B ;  v^+ = (R x t + (R.t) t + R) / (1 + t.t)
  const R t[3] = {h * B[0], h * B[1], h * B[2]};
  const R Rv[3] = {vm[0] + (vm[1] * t[2] - vm[2] * t[1]), vm[1] + (vm[2] * t[0] - vm[0] * t[2]), vm[2] + (vm[0] * t[1] - vm[1] * t[0])};
  const R Rt = Rv[0] * t[0] + Rv[1] * t[1] + Rv[2] * t[2];
  const R inv = R(1) / (R(1) + t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
  const R vp[3] = {(Rv[1] * t[2] - Rv[2] * t[1] + Rt * t[0] + Rv[0]) * inv, (Rv[2] * t[0] - Rv[0] * t[2] + Rt * t[1] + Rv[1]) * inv,
                   (Rv[0] * t[1] - Rv[1] * t[0] + Rt * t[2] + Rv[2]) * inv};
  v[0] = vp[0] + h * E[0];
  v[1] = vp[1] + h * E[1];
  v[2] = vp[2] + h * E[2];
}

template <typename R>
__device__ __forceinline__ void boris_velocity_relativistic(R v[3], const R E[3], const R B[3], R q, R m, R dt) {
  // the arithmetic runs in double even for R = float: m^2 c^2 of a weighted macro-particle underflows fp32
  const double c = kC;
  const double vx = v[0], vy = v[1], vz = v[2];
  const double gamma_n = 1.0 / sqrt(1.0 - (vx * vx + vy * vy + vz * vz) / (c * c));
  const double qd = q, md = m, h = (double)q * (double)dt * 0.5;
  double pm[3] = {gamma_n * md * vx + h * E[0], gamma_n * md * vy + h * E[1], gamma_n * md * vz + h * E[2]};
  const double gamma_m = sqrt(1.0 + (pm[0] * pm[0] + pm[1] * pm[1] + pm[2] * pm[2]) / (md * md * c * c));
  const double f = (qd * (double)dt) / (2.0 * md * gamma_m);
  const double t[3] = {f * B[0], f * B[1], f * B[2]};
  const double pdt = pm[0] * t[0] + pm[1] * t[1] + pm[2] * t[2];
  const double t2 = t[0] * t[0] + t[1] * t[1] + t[2] * t[2];
  const double px[3] = {pm[1] * t[2] - pm[2] * t[1], pm[2] * t[0] - pm[0] * t[2], pm[0] * t[1] - pm[1] * t[0]};
  double pn[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) pn[a] = (pm[a] * (1.0 - t2) + 2.0 * (pdt * t[a] + px[a])) / (1.0 + t2) + h * E[a];
  const double mc = md * c;
  const double gamma_new = sqrt(1.0 + (pn[0] * pn[0] + pn[1] * pn[1] + pn[2] * pn[2]) / (mc * mc));
#pragma unroll
  for (int a = 0; a < 3; ++a) v[a] = (R)(pn[a] / (gamma_new * md));
}

template <typename R>
__device__ __forceinline__ int species_of(long long i, const DevParams<R>& p) {
  int s = 0;
  while (s < p.n_species - 1 && i >= p.sp_end[s]) ++s;
  return s;
}

}  // namespace jic
