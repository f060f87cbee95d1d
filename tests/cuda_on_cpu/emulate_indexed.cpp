// TEST INFRASTRUCTURE ONLY -- the INDEXED engine's kernels compiled as host code (see fake_cuda/cuda_runtime.h) and driven the way
// csrc/jic_engine.cu drives them:  k_start -> k_fields(init) -> T x ( k_step<double,false> -> k_fields ), with the option to stop after
// `reload_at` steps, take the reference-shaped carry out of the state and put it back through k_load_carry -> k_fields(init) ->
// k_carry_fields (what jic_load_carry does) before continuing.  tests/test_cuda_source_on_cpu.py compares the histories with the oracle.
#include <cstring>
#include <vector>

#include "jic_kernels.cuh"
#include "jic_carry.cuh"

EmuDim3 threadIdx, blockIdx;
EmuDim3One blockDim, gridDim;
namespace jic {
alignas(16) double fsm[1];            // k_fields with smem_comps = 0 never touches it
alignas(16) unsigned char smem_raw[1];
alignas(16) double gsm[1];
}  // namespace jic

using namespace jic;

extern "C" {

struct EmuParams {
  int G, n_species, pbl, pbr, fbl, fbr, relativistic, field_solver_unused;
  int filter_passes, n_strides, strides[8];
  double L, Ly, Lz, dx, dt, grid_first, grid_last, filter_alpha;
  long long count[8];
  double q[8], m[8], qm[8];
};

}  // extern "C"

template <typename R>
static int run(const EmuParams* ep, const R* x0, const R* v0, const float* extE_f, const float* extB_f, int T, int reload_at,
               R* hE, R* hB, R* hJ, R* hrho, R* hx, R* hv, double* E0_out, R* vinit_out) {
  DevParams<R> p;
  std::memset(&p, 0, sizeof(p));
  long long N = 0;
  for (int s = 0; s < ep->n_species; ++s) { N += ep->count[s]; p.sp_end[s] = N; p.sp_q[s] = ep->q[s]; p.sp_m[s] = ep->m[s]; p.sp_qm[s] = ep->qm[s]; }
  p.N = N; p.G = ep->G; p.n_species = ep->n_species;
  p.pbl = ep->pbl; p.pbr = ep->pbr; p.fbl = ep->fbl; p.fbr = ep->fbr; p.relativistic = ep->relativistic; p.track_yz = 1; p.stag = 0;
  // exactly the assignments of EngineT::create (csrc/jic_engine.cu)
  p.L = ep->L; p.Ly = ep->Ly; p.Lz = ep->Lz; p.half_L = ep->L / 2; p.half_Ly = ep->Ly / 2; p.half_Lz = ep->Lz / 2;
  p.dx = ep->dx; p.inv_dx = 1.0 / ep->dx; p.half_dx = ep->dx / 2; p.dt = ep->dt; p.half_dt = ep->dt / 2;
  p.g0 = ep->grid_first; p.gl = ep->grid_last; p.gs = ep->grid_first - ep->dx / 2;
  p.park_left = ep->grid_first - 1.5 * ep->dx; p.park_right = ep->grid_last + 3 * ep->dx;
  p.park_left_cell = reference_floor_div((ep->grid_first - 1.5 * ep->dx) - (ep->grid_first - ep->dx / 2), ep->dx);
  const size_t G = (size_t)ep->G, n = (size_t)N;
  std::vector<R> xh(n), yh(n), zh(n), vx(n), vy(n), vz(n), v_init(3 * n), acc(G * (kAccRow + 1)), F((G + 3) * kFieldRow);
  std::vector<double> E(G * 3), B(G * 3), E_int(G * 3), B_int(G * 3), J(G * 3), rho(G), extE(G * 3), extB(G * 3), s0(G * kAccRow), s1(G * kAccRow),
      E0(G * 3), B0(G * 3);
  for (size_t k = 0; k < G * 3; ++k) { extE[k] = extE_f ? (double)extE_f[k] : 0.0; extB[k] = extB_f ? (double)extB_f[k] : 0.0; }
  RunControl ctl;
  std::memset(&ctl, 0, sizeof(ctl));
  auto field_args = [&](bool init) {  // EngineT::field_args
    FieldArgs<R> a;
    std::memset(&a, 0, sizeof(a));
    a.G = ep->G; a.fbl = ep->fbl; a.fbr = ep->fbr; a.passes = ep->filter_passes; a.n_strides = ep->n_strides; a.init = init;
    for (int i = 0; i < ep->n_strides; ++i) a.strides[i] = ep->strides[i];
    a.alpha = ep->filter_alpha; a.dx = ep->dx; a.dt = ep->dt;
    a.acc = acc.data(); a.E = E.data(); a.B = B.data(); a.E_int = E_int.data(); a.B_int = B_int.data(); a.J = J.data(); a.rho = rho.data();
    a.extE = extE.data(); a.extB = extB.data(); a.F = F.data(); a.s0 = s0.data(); a.s1 = s1.data(); a.E0 = E0.data(); a.B0 = B0.data(); a.ctl = &ctl;
    a.record = init ? 0 : 1; a.smem_comps = 0; a.ExC = nullptr;
    return a;
  };
  // jic_initialize
  k_start<R>(p, x0, v0, 0, N, xh.data(), yh.data(), zh.data(), vx.data(), vy.data(), vz.data(), v_init.data(), acc.data());
  k_fields<R>(field_args(true));
  if (E0_out) std::memcpy(E0_out, E0.data(), sizeof(double) * G * 3);
  if (vinit_out) std::memcpy(vinit_out, v_init.data(), sizeof(R) * 3 * n);
  std::vector<R> x_minus(3 * n), x_plus(3 * n), x_now(3 * n), v_now(3 * n);
  for (int t = 0; t < T; ++t) {
    if (t == reload_at && t > 0) {
      // the reference-shaped carry after t steps: (E, B) at integer time, x_{n-1/2} = the x_{n+1/2} of the step before, x_n, x_{n+1/2}, v_n
      for (size_t i = 0; i < n; ++i) {
        x_plus[3 * i] = xh[i]; x_plus[3 * i + 1] = yh[i]; x_plus[3 * i + 2] = zh[i];
        v_now[3 * i] = vx[i]; v_now[3 * i + 1] = vy[i]; v_now[3 * i + 2] = vz[i];
        for (int c = 0; c < 3; ++c) x_now[3 * i + c] = hx[((size_t)(t - 1) * n + i) * 3 + c];
      }
      std::vector<R> E_c(E_int.begin(), E_int.end()), B_c(B_int.begin(), B_int.end());  // (the carry travels in R, as through the C ABI)
      // wipe the state that jic_load_carry must rebuild
      std::fill(xh.begin(), xh.end(), 0.0); std::fill(vx.begin(), vx.end(), 0.0); std::fill(vy.begin(), vy.end(), 0.0); std::fill(vz.begin(), vz.end(), 0.0);
      std::fill(E.begin(), E.end(), 1e300); std::fill(B.begin(), B.end(), 1e300); std::fill(J.begin(), J.end(), 1e300); std::fill(F.begin(), F.end(), R(1e30));
      std::fill(acc.begin(), acc.end(), 0.0);
      const long long row = ctl.hist_row, step = ctl.step;
      k_load_carry<R>(p, x_minus.data(), x_now.data(), x_plus.data(), v_now.data(), xh.data(), yh.data(), zh.data(), vx.data(), vy.data(), vz.data(),
                      v_init.data(), acc.data());
      k_fields<R>(field_args(true));
      k_carry_fields<R>(field_args(true), E_c.data(), B_c.data());
      ctl.hist_row = row; ctl.step = step;  // (the library restarts its histories at row 0 of a new jic_run; the harness keeps one buffer)
    }
    for (size_t i = 0; i < n; ++i) { x_minus[3 * i] = xh[i]; x_minus[3 * i + 1] = yh[i]; x_minus[3 * i + 2] = zh[i]; }
    ctl.hist[0] = hE; ctl.hist[1] = hB; ctl.hist[2] = hJ; ctl.hist[3] = hrho; ctl.hist[4] = hx; ctl.hist[5] = hv;
    k_step<R, false>(p, xh.data(), yh.data(), zh.data(), vx.data(), vy.data(), vz.data(), F.data(), acc.data(), &ctl);
    k_fields<R>(field_args(false));
  }
  return 0;
}

extern "C" {
__attribute__((visibility("default"))) int emu_run(const EmuParams* ep, const double* x0, const double* v0, const float* extE, const float* extB, int T,
                                                   int reload_at, double* hE, double* hB, double* hJ, double* hrho, double* hx, double* hv,
                                                   double* E0_out, double* vinit_out) {
  return run<double>(ep, x0, v0, extE, extB, T, reload_at, hE, hB, hJ, hrho, hx, hv, E0_out, vinit_out);
}
__attribute__((visibility("default"))) int emu_run_f32(const EmuParams* ep, const float* x0, const float* v0, const float* extE, const float* extB, int T,
                                                       int reload_at, float* hE, float* hB, float* hJ, float* hrho, float* hx, float* hv,
                                                       double* E0_out, float* vinit_out) {
  return run<float>(ep, x0, v0, extE, extB, T, reload_at, hE, hB, hJ, hrho, hx, hv, E0_out, vinit_out);
}
}  // extern "C"
