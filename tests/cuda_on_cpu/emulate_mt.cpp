// TEST INFRASTRUCTURE ONLY -- the warp-level kernels of the INDEXED engine and of the Crank-Nicolson stepper on the multi-threaded
// emulation of fake_cuda_mt/cuda_runtime.h, launched in the order csrc/jic_engine.cu launches them:
//   emu_fs_run   explicit stepper with the per-step electrostatic correction: k_start, k_gauss_kernel, { k_step (face deposit), k_gauss,
//                k_fields (E_x replaced) } -- optional carry reload through k_load_carry / k_carry_fields
//   emu_cn_run   implicit stepper: k_cn_start, k_fields(init), k_cn_fields(prepare), { max_iter x (k_cn_push, k_cn_fields), k_cn_record }
//                -- optional carry reload through k_cn_load / k_carry_copy_fields
//   emu_cn_sorted_run   the same with the cell-sorted push of csrc/jic_cn_sorted.cuh, as EngineT::enqueue_step_cn_sorted launches it:
//                k_cn_meta_init, { k_cn_hist, k_cn_scan, k_cn_scatter, max_iter x (k_cn_push_sorted, k_cn_fields), k_cn_record_sorted };
//                the reload goes through k_cn_export_sorted (what jic_get_particles runs)
#include <cstring>
#include <memory>
#include <vector>

#include "jic_kernels.cuh"
#include "jic_cn.cuh"
#include "jic_cn_sorted.cuh"
#include "jic_carry.cuh"

thread_local EmuDim3 threadIdx;
EmuDim3 blockIdx;
EmuDim3One blockDim, gridDim;
EmuBlock* emu_block = nullptr;
namespace jic {
alignas(16) double fsm[1];
alignas(16) unsigned char smem_raw[1];
alignas(16) unsigned char cn_smem_raw[1];
alignas(256) unsigned char cn_sorted_smem[(kCnSortedThreads / 32) * kCnWin * kCnWinComps * kCnColStride * sizeof(double)];
alignas(16) double gsm[2 * 4096];
}  // namespace jic

using namespace jic;
typedef double R;

extern "C" {
struct EmuParams {
  int G, n_species, pbl, pbr, fbl, fbr, relativistic, field_solver;
  int filter_passes, n_strides, strides[8];
  double L, Ly, Lz, dx, dt, grid_first, grid_last, filter_alpha;
  long long count[8];
  double q[8], m[8], qm[8];
};
}

static DevParams<R> dev_params(const EmuParams* ep, long long* N_out) {
  DevParams<R> p;
  std::memset(&p, 0, sizeof(p));
  long long N = 0;
  for (int s = 0; s < ep->n_species; ++s) { N += ep->count[s]; p.sp_end[s] = N; p.sp_q[s] = ep->q[s]; p.sp_m[s] = ep->m[s]; p.sp_qm[s] = ep->qm[s]; }
  p.N = N; p.G = ep->G; p.n_species = ep->n_species;
  p.pbl = ep->pbl; p.pbr = ep->pbr; p.fbl = ep->fbl; p.fbr = ep->fbr; p.relativistic = ep->relativistic; p.track_yz = 1;
  p.L = ep->L; p.Ly = ep->Ly; p.Lz = ep->Lz; p.half_L = ep->L / 2; p.half_Ly = ep->Ly / 2; p.half_Lz = ep->Lz / 2;
  p.dx = ep->dx; p.inv_dx = 1.0 / ep->dx; p.half_dx = ep->dx / 2; p.dt = ep->dt; p.half_dt = ep->dt / 2;
  p.g0 = ep->grid_first; p.gl = ep->grid_last; p.gs = ep->grid_first - ep->dx / 2;
  p.park_left = ep->grid_first - 1.5 * ep->dx; p.park_right = ep->grid_last + 3 * ep->dx;
  p.park_left_cell = reference_floor_div((ep->grid_first - 1.5 * ep->dx) - (ep->grid_first - ep->dx / 2), ep->dx);
  *N_out = N;
  return p;
}

struct GridState {
  std::vector<R> acc, F;
  std::vector<double> E, B, E_int, B_int, J, rho, extE, extB, s0, s1, E0, B0, ExC, h;
  RunControl ctl;
  explicit GridState(size_t G) : acc(G * (kAccRow + 1)), F((G + 3) * kFieldRow), E(G * 3), B(G * 3), E_int(G * 3), B_int(G * 3), J(G * 3), rho(G),
                                 extE(G * 3), extB(G * 3), s0(G * kAccRow), s1(G * kAccRow), E0(G * 3), B0(G * 3), ExC(G), h(G) {
    std::memset(&ctl, 0, sizeof(ctl));
  }
  FieldArgs<R> field_args(const EmuParams* ep, bool init, bool with_ExC) {
    FieldArgs<R> a;
    std::memset(&a, 0, sizeof(a));
    a.G = ep->G; a.fbl = ep->fbl; a.fbr = ep->fbr; a.passes = ep->filter_passes; a.n_strides = ep->n_strides; a.init = init;
    for (int i = 0; i < ep->n_strides; ++i) a.strides[i] = ep->strides[i];
    a.alpha = ep->filter_alpha; a.dx = ep->dx; a.dt = ep->dt;
    a.acc = acc.data(); a.E = E.data(); a.B = B.data(); a.E_int = E_int.data(); a.B_int = B_int.data(); a.J = J.data(); a.rho = rho.data();
    a.extE = extE.data(); a.extB = extB.data(); a.F = F.data(); a.s0 = s0.data(); a.s1 = s1.data(); a.E0 = E0.data(); a.B0 = B0.data(); a.ctl = &ctl;
    a.record = init ? 0 : 1; a.smem_comps = 0; a.ExC = (with_ExC && !init) ? ExC.data() : nullptr;
    return a;
  }
};

constexpr unsigned kThreads = 64;  // two warps per block: block reductions and warp loops both have something to do

extern "C" {

}  // extern "C"

// n_chunks > 1: the start-up kernels run chunk by chunk on chunk-local x0/v0 pointers with the global offset i0, as EngineT::initialize_host
// enqueues them for a pipelined host upload
static int fs_run(const EmuParams* ep, const double* x0, const double* v0, int T, int reload_at, int n_chunks, double* hE, double* hB, double* hJ,
                  double* hrho, double* hx, double* hv) {
  long long N;
  DevParams<R> p = dev_params(ep, &N);
  p.stag = ep->field_solver != 0;
  const size_t G = (size_t)ep->G, n = (size_t)N;
  if (2 * G > sizeof(gsm) / sizeof(double)) return -1;
  GridState gs(G);
  std::vector<R> xh(n), yh(n), zh(n), vx(n), vy(n), vz(n), v_init(3 * n), x_minus(3 * n), x_plus(3 * n), x_now(3 * n), v_now(3 * n);
  const bool stag = p.stag != 0;
  if (stag) emu_launch((unsigned)((G + 127) / 128), 128, [&] { k_gauss_kernel((int)G, ep->dx, gs.h.data()); });
  const bool fix = stag && ep->pbl != ep->pbr && (ep->pbl == JIC_BC_PERIODIC || ep->pbr == JIC_BC_PERIODIC);  // EngineT::needs_face_fix
  for (int c = 0; c < n_chunks; ++c) {
    const long long i0 = N * c / n_chunks, cn = N * (c + 1) / n_chunks - i0;
    std::vector<double> cx(x0 + 3 * i0, x0 + 3 * (i0 + cn)), cv(v0 + 3 * i0, v0 + 3 * (i0 + cn));  // a staging buffer: only this chunk is addressable
    emu_launch(2, kThreads, [&] { k_start<R>(p, cx.data(), cv.data(), i0, cn, xh.data(), yh.data(), zh.data(), vx.data(), vy.data(), vz.data(), v_init.data(),
                                             gs.acc.data()); });
    if (fix && cn > 0) emu_launch(2, kThreads, [&] { k_start_face_fix<R>(p, cx.data(), cv.data(), i0, cn, gs.acc.data()); });
  }
  emu_launch(1, kThreads, [&] { k_fields<R>(gs.field_args(ep, true, stag)); });
  for (int t = 0; t < T; ++t) {
    if (t == reload_at && t > 0) {
      for (size_t i = 0; i < n; ++i) {
        x_plus[3 * i] = xh[i]; x_plus[3 * i + 1] = yh[i]; x_plus[3 * i + 2] = zh[i];
        v_now[3 * i] = vx[i]; v_now[3 * i + 1] = vy[i]; v_now[3 * i + 2] = vz[i];
        for (int c = 0; c < 3; ++c) x_now[3 * i + c] = hx[((size_t)(t - 1) * n + i) * 3 + c];
      }
      std::vector<R> E_c(gs.E_int.begin(), gs.E_int.end()), B_c(gs.B_int.begin(), gs.B_int.end());
      std::fill(xh.begin(), xh.end(), 0.0); std::fill(vx.begin(), vx.end(), 0.0); std::fill(vy.begin(), vy.end(), 0.0); std::fill(vz.begin(), vz.end(), 0.0);
      std::fill(gs.E.begin(), gs.E.end(), 1e300); std::fill(gs.B.begin(), gs.B.end(), 1e300); std::fill(gs.J.begin(), gs.J.end(), 1e300);
      std::fill(gs.F.begin(), gs.F.end(), 1e300); std::fill(gs.acc.begin(), gs.acc.end(), 0.0);
      const long long row = gs.ctl.hist_row, step = gs.ctl.step;
      emu_launch(2, kThreads, [&] { k_load_carry<R>(p, x_minus.data(), x_now.data(), x_plus.data(), v_now.data(), xh.data(), yh.data(), zh.data(), vx.data(),
                                                     vy.data(), vz.data(), v_init.data(), gs.acc.data()); });
      emu_launch(1, kThreads, [&] { k_fields<R>(gs.field_args(ep, true, stag)); });
      emu_launch(1, kThreads, [&] { k_carry_fields<R>(gs.field_args(ep, true, stag), E_c.data(), B_c.data()); });
      gs.ctl.hist_row = row; gs.ctl.step = step;
    }
    for (size_t i = 0; i < n; ++i) { x_minus[3 * i] = xh[i]; x_minus[3 * i + 1] = yh[i]; x_minus[3 * i + 2] = zh[i]; }
    gs.ctl.hist[0] = hE; gs.ctl.hist[1] = hB; gs.ctl.hist[2] = hJ; gs.ctl.hist[3] = hrho; gs.ctl.hist[4] = hx; gs.ctl.hist[5] = hv;
    emu_launch(3, kThreads, [&] { k_step<R, false>(p, xh.data(), yh.data(), zh.data(), vx.data(), vy.data(), vz.data(), gs.F.data(), gs.acc.data(), &gs.ctl); });
    if (stag) {  // EngineT::enqueue_gauss
      GaussArgs<R> a;
      std::memset(&a, 0, sizeof(a));
      a.G = ep->G; a.fbl = ep->fbl; a.fbr = ep->fbr; a.passes = ep->filter_passes; a.n_strides = ep->n_strides; a.mode = ep->field_solver;
      for (int i = 0; i < ep->n_strides; ++i) a.strides[i] = ep->strides[i];
      a.alpha = ep->filter_alpha; a.dx = ep->dx; a.accS = gs.acc.data() + G * kAccRow; a.h = gs.h.data(); a.Ex = gs.ExC.data();
      emu_launch((unsigned)((G + 7) / 8), kThreads, [&] { k_gauss<R>(a); });
    }
    emu_launch(1, kThreads, [&] { k_fields<R>(gs.field_args(ep, false, stag)); });
  }
  return 0;
}

extern "C" {

__attribute__((visibility("default"))) int emu_fs_run(const EmuParams* ep, const double* x0, const double* v0, int T, int reload_at, double* hE, double* hB,
                                                      double* hJ, double* hrho, double* hx, double* hv) {
  return fs_run(ep, x0, v0, T, reload_at, 1, hE, hB, hJ, hrho, hx, hv);
}
__attribute__((visibility("default"))) int emu_fs_run_chunked(const EmuParams* ep, const double* x0, const double* v0, int T, int n_chunks, double* hE,
                                                              double* hB, double* hJ, double* hrho, double* hx, double* hv) {
  return fs_run(ep, x0, v0, T, -1, n_chunks, hE, hB, hJ, hrho, hx, hv);
}

// Two ranks of the explicit stepper with the per-step electrostatic correction, reduced the way EngineT does it through NCCL: every
// rank deposits its own particles into its own raw grid, the grids (face component included) are summed before the field kernels, and
// after the START-UP reduction the face component survives on rank 0 only (EngineT::initialize_finish) -- it holds the step-0
// correction of k_start_face_fix, which the start-up field kernel does not consume and which step 0's reduction must count once.
// Field histories of rank 0 are returned; both ranks must end up with identical fields (checked here, -2 otherwise).
__attribute__((visibility("default"))) int emu_fs_run_two_ranks(const EmuParams* ep0, const EmuParams* ep1, const double* x0_0, const double* v0_0,
                                                                const double* x0_1, const double* v0_1, int T, int keep_on_all_ranks, double* hE,
                                                                double* hB, double* hJ, double* hrho) {
  const EmuParams* eps[2] = {ep0, ep1};
  const double* x0s[2] = {x0_0, x0_1};
  const double* v0s[2] = {v0_0, v0_1};
  const size_t G = (size_t)ep0->G;
  if (2 * G > sizeof(gsm) / sizeof(double) || ep0->field_solver == 0) return -1;
  struct Rank {
    DevParams<R> p;
    long long N = 0;
    std::unique_ptr<GridState> gs;
    std::vector<R> xh, yh, zh, vx, vy, vz, v_init;
    std::vector<double> E, B, J, rho;
  } rk[2];
  const size_t n_red = G * (kAccRow + 1);
  auto allreduce = [&] {
    for (size_t k = 0; k < n_red; ++k) { const R s = rk[0].gs->acc[k] + rk[1].gs->acc[k]; rk[0].gs->acc[k] = s; rk[1].gs->acc[k] = s; }
  };
  const bool fix = ep0->pbl != ep0->pbr && (ep0->pbl == JIC_BC_PERIODIC || ep0->pbr == JIC_BC_PERIODIC);
  for (int r = 0; r < 2; ++r) {
    Rank& k = rk[r];
    k.p = dev_params(eps[r], &k.N);
    k.p.stag = 1;
    k.gs = std::make_unique<GridState>(G);
    const size_t n = (size_t)k.N;
    k.xh.resize(n); k.yh.resize(n); k.zh.resize(n); k.vx.resize(n); k.vy.resize(n); k.vz.resize(n); k.v_init.resize(3 * n);
    k.E.resize((size_t)T * G * 3); k.B.resize((size_t)T * G * 3); k.J.resize((size_t)T * G * 3); k.rho.resize((size_t)T * G);
    emu_launch((unsigned)((G + 127) / 128), 128, [&] { k_gauss_kernel((int)G, eps[r]->dx, k.gs->h.data()); });
    emu_launch(2, kThreads, [&] { k_start<R>(k.p, x0s[r], v0s[r], 0, k.N, k.xh.data(), k.yh.data(), k.zh.data(), k.vx.data(), k.vy.data(), k.vz.data(),
                                             k.v_init.data(), k.gs->acc.data()); });
    if (fix) emu_launch(2, kThreads, [&] { k_start_face_fix<R>(k.p, x0s[r], v0s[r], 0, k.N, k.gs->acc.data()); });
  }
  allreduce();
  if (fix && !keep_on_all_ranks) std::fill(rk[1].gs->acc.begin() + G * kAccRow, rk[1].gs->acc.end(), R(0));
  for (int r = 0; r < 2; ++r) emu_launch(1, kThreads, [&] { k_fields<R>(rk[r].gs->field_args(eps[r], true, true)); });
  for (int t = 0; t < T; ++t) {
    for (int r = 0; r < 2; ++r) {
      Rank& k = rk[r];
      k.gs->ctl.hist[0] = k.E.data(); k.gs->ctl.hist[1] = k.B.data(); k.gs->ctl.hist[2] = k.J.data(); k.gs->ctl.hist[3] = k.rho.data();
      k.gs->ctl.hist[4] = nullptr; k.gs->ctl.hist[5] = nullptr;
      emu_launch(3, kThreads, [&] { k_step<R, false>(k.p, k.xh.data(), k.yh.data(), k.zh.data(), k.vx.data(), k.vy.data(), k.vz.data(), k.gs->F.data(),
                                                     k.gs->acc.data(), &k.gs->ctl); });
    }
    allreduce();
    for (int r = 0; r < 2; ++r) {
      Rank& k = rk[r];
      GaussArgs<R> a;
      std::memset(&a, 0, sizeof(a));
      a.G = ep0->G; a.fbl = ep0->fbl; a.fbr = ep0->fbr; a.passes = ep0->filter_passes; a.n_strides = ep0->n_strides; a.mode = ep0->field_solver;
      for (int i = 0; i < ep0->n_strides; ++i) a.strides[i] = ep0->strides[i];
      a.alpha = ep0->filter_alpha; a.dx = ep0->dx; a.accS = k.gs->acc.data() + G * kAccRow; a.h = k.gs->h.data(); a.Ex = k.gs->ExC.data();
      emu_launch((unsigned)((G + 7) / 8), kThreads, [&] { k_gauss<R>(a); });
      emu_launch(1, kThreads, [&] { k_fields<R>(k.gs->field_args(eps[r], false, true)); });
    }
  }
  if (rk[0].E != rk[1].E || rk[0].B != rk[1].B || rk[0].J != rk[1].J || rk[0].rho != rk[1].rho) return -2;
  std::memcpy(hE, rk[0].E.data(), rk[0].E.size() * sizeof(double)); std::memcpy(hB, rk[0].B.data(), rk[0].B.size() * sizeof(double));
  std::memcpy(hJ, rk[0].J.data(), rk[0].J.size() * sizeof(double)); std::memcpy(hrho, rk[0].rho.data(), rk[0].rho.size() * sizeof(double));
  return 0;
}

static int g_cn_sort_every = 1;  // EngineT::cn_sort_every (JIC_CN_SORT_EVERY)
extern "C" __attribute__((visibility("default"))) void emu_set_cn_sort_every(int k) { g_cn_sort_every = k < 1 ? 1 : k; }

static int emu_cn_run_impl(bool sorted, const EmuParams* ep, const double* x0, const double* v0, int T, int n_sub, int max_iter, double tol,
                           int reload_at, double* hE, double* hB, double* hJ, double* hrho, double* hx, double* hv, long long* picard) {
  long long N;
  DevParams<R> p = dev_params(ep, &N);
  const size_t G = (size_t)ep->G, n = (size_t)N;
  GridState gs(G);
  std::vector<R> buf[2][6], stag(n * (size_t)n_sub), v_init(3 * n);
  CnState<R> cs[2];
  for (int k = 0; k < 2; ++k) {
    for (auto& b : buf[k]) b.assign(n, 0.0);
    cs[k] = CnState<R>{buf[k][0].data(), buf[k][1].data(), buf[k][2].data(), buf[k][3].data(), buf[k][4].data(), buf[k][5].data()};
  }
  std::vector<uint8_t> alive(n), alive2(n), sp(n), sp2(n);
  std::vector<int> perm(n), perm2(n);
  std::vector<unsigned> hist(G, 0u), off(G, 0u);
  std::vector<double> Eg(G * 3), Bnext(G * 3), Eavg(G * 3), Bavg(G * 3);
  std::vector<double> EB_store(G * 6 + 2);
  double* EB = sorted ? (double*)(((uintptr_t)EB_store.data() + 15) & ~(uintptr_t)15) : nullptr;  // (read 16 bytes at a time)
  CnControl cn;
  std::memset(&cn, 0, sizeof(cn));
  auto cn_args = [&](int it, bool prepare_only) {  // EngineT::cn_field_args
    CnFieldArgs<R> a;
    std::memset(&a, 0, sizeof(a));
    a.G = ep->G; a.fbl = ep->fbl; a.fbr = ep->fbr; a.it = it; a.max_iter = max_iter; a.prepare_only = prepare_only ? 1 : 0;
    a.dx = ep->dx; a.dt = ep->dt; a.tol = tol;
    a.acc = gs.acc.data(); a.En = gs.E.data(); a.Bn = gs.B.data(); a.Eg = Eg.data(); a.Bnext = Bnext.data(); a.Eavg = Eavg.data(); a.Bavg = Bavg.data();
    a.J = gs.J.data(); a.rho = gs.rho.data(); a.cn = &cn; a.ctl = &gs.ctl; a.EB = EB;
    return a;
  };
  // EngineT::initialize_cn
  emu_launch(2, kThreads, [&] { k_cn_start<R>(p, x0, v0, cs[0], v_init.data(), alive.data(), gs.acc.data()); });
  if (sorted) emu_launch(2, kThreads, [&] { k_cn_meta_init<R>(p, perm.data(), sp.data()); });
  emu_launch(1, kThreads, [&] { k_fields<R>(gs.field_args(ep, true, false)); });
  gs.E = gs.E0; gs.B = gs.B0;
  emu_launch(1, kThreads, [&] { k_cn_fields<R>(cn_args(0, true)); });
  int par = 0, age = 0;
  for (int t = 0; t < T; ++t) {
    if (t == reload_at && t > 0) {
      std::vector<R> x(3 * n), v(3 * n), E_c(gs.E.begin(), gs.E.end()), B_c(gs.B.begin(), gs.B.end());
      std::vector<uint8_t> alive_in(alive);
      if (sorted) {  // EngineT::get_particles; the q = 0 bytes back in input order
        emu_launch(2, kThreads, [&] { k_cn_export_sorted<R>(p, cs[par], perm.data(), x.data(), v.data(), (uint8_t*)nullptr); });
        for (size_t i = 0; i < n; ++i) alive_in[(size_t)perm[i]] = alive[i];
        std::fill(perm.begin(), perm.end(), -1); std::fill(sp.begin(), sp.end(), 99);
      } else {
        for (size_t i = 0; i < n; ++i) {
          x[3 * i] = cs[par].x[i]; x[3 * i + 1] = cs[par].y[i]; x[3 * i + 2] = cs[par].z[i];
          v[3 * i] = cs[par].vx[i]; v[3 * i + 1] = cs[par].vy[i]; v[3 * i + 2] = cs[par].vz[i];
        }
      }
      for (int k = 0; k < 2; ++k) for (auto& b : buf[k]) std::fill(b.begin(), b.end(), 1e300);
      std::fill(alive.begin(), alive.end(), 7); std::fill(gs.E.begin(), gs.E.end(), 1e300); std::fill(gs.B.begin(), gs.B.end(), 1e300);
      std::fill(Eg.begin(), Eg.end(), 1e300); std::fill(Eavg.begin(), Eavg.end(), 1e300); std::fill(Bavg.begin(), Bavg.end(), 1e300);
      std::fill(EB_store.begin(), EB_store.end(), 1e300);
      std::fill(gs.acc.begin(), gs.acc.end(), 0.0);
      const CnControl keep = cn;
      std::memset(&cn, 0, sizeof(cn));
      par = 0; age = 0;  // EngineT::load_carry_cn
      emu_launch(2, kThreads, [&] { k_cn_load<R>(p, x.data(), v.data(), alive_in.data(), cs[0], v_init.data(), alive.data()); });
      if (sorted) emu_launch(2, kThreads, [&] { k_cn_meta_init<R>(p, perm.data(), sp.data()); });
      emu_launch(1, kThreads, [&] { k_carry_copy_fields<R>(E_c.data(), B_c.data(), gs.E.data(), gs.B.data(), gs.E0.data(), gs.B0.data(), (int)(G * 3)); });
      emu_launch(1, kThreads, [&] { k_cn_fields<R>(cn_args(0, true)); });
      cn.total_iters = keep.total_iters;
    }
    gs.ctl.hist[0] = hE; gs.ctl.hist[1] = hB; gs.ctl.hist[2] = hJ; gs.ctl.hist[3] = hrho; gs.ctl.hist[4] = hx; gs.ctl.hist[5] = hv;
    if (sorted) {  // EngineT::enqueue_step_cn_sorted + cn_sorted_advance
      int src = par, dst = par ^ 1;
      if (age == 0) {
        emu_launch(3, kThreads, [&] { k_cn_hist<R>(p, cs[par].x, hist.data()); });
        emu_launch(1, kThreads, [&] { k_cn_scan((int)G, hist.data(), off.data()); });
        emu_launch(3, kThreads, [&] { k_cn_scatter<R>(p, cs[par], cs[par ^ 1], perm.data(), perm2.data(), sp.data(), sp2.data(), alive.data(), alive2.data(),
                                                      off.data(), hist.data()); });
        std::fill(hist.begin(), hist.end(), 0u);
        perm = perm2; sp = sp2; alive = alive2;
        src = par ^ 1; dst = par;
      }
      for (int it = 0; it < max_iter; ++it) {
        emu_launch(3, kThreads, [&] { k_cn_push_sorted<R>(p, cs[src], cs[dst], stag.data(), n_sub, it, EB, gs.acc.data(),
                                                          alive.data(), sp.data(), &cn); });
        emu_launch(1, kThreads, [&] { k_cn_fields<R>(cn_args(it, false)); });
      }
      emu_launch(2, kThreads, [&] { k_cn_record_sorted<R>(p, cs[dst], perm.data(), &gs.ctl); });
      if (picard) picard[t] = cn.last_iters;
      if (age != 0) par ^= 1;
      age = (age + 1) % g_cn_sort_every;
      continue;
    }
    for (int it = 0; it < max_iter; ++it) {  // EngineT::enqueue_step_cn
      emu_launch(3, kThreads, [&] { k_cn_push<R, false>(p, cs[par], cs[par ^ 1], stag.data(), n_sub, it, Eavg.data(), Bavg.data(), gs.acc.data(), alive.data(), &cn); });
      emu_launch(1, kThreads, [&] { k_cn_fields<R>(cn_args(it, false)); });
    }
    emu_launch(2, kThreads, [&] { k_cn_record<R>(p, cs[par ^ 1], &gs.ctl); });
    if (picard) picard[t] = cn.last_iters;
    par ^= 1;
  }
  return 0;
}

__attribute__((visibility("default"))) int emu_cn_run(const EmuParams* ep, const double* x0, const double* v0, int T, int n_sub, int max_iter, double tol,
                                                      int reload_at, double* hE, double* hB, double* hJ, double* hrho, double* hx, double* hv,
                                                      long long* picard) {
  return emu_cn_run_impl(false, ep, x0, v0, T, n_sub, max_iter, tol, reload_at, hE, hB, hJ, hrho, hx, hv, picard);
}
__attribute__((visibility("default"))) int emu_cn_sorted_run(const EmuParams* ep, const double* x0, const double* v0, int T, int n_sub, int max_iter,
                                                             double tol, int reload_at, double* hE, double* hB, double* hJ, double* hrho, double* hx,
                                                             double* hv, long long* picard) {
  return emu_cn_run_impl(true, ep, x0, v0, T, n_sub, max_iter, tol, reload_at, hE, hB, hJ, hrho, hx, hv, picard);
}

}  // extern "C"
