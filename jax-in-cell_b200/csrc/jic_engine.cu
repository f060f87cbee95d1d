// Host side of libjic_b200: contexts, CUDA-graph time loop, NCCL plumbing and the C ABI of include/jic_b200.h.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <unistd.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "jic_host.cuh"
#include "jic_kernels.cuh"
#include "jic_binned.cuh"
#include "jic_sample.cuh"
#include "jic_cn.cuh"
#include "jic_cn_sorted.cuh"
#include "jic_carry.cuh"

namespace jic {

// ---------------------------------------------------------------------------------------------------------
// NCCL, resolved at run time so that the library loads (and single-GPU runs work) without it.
// ---------------------------------------------------------------------------------------------------------
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};

static NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* env = getenv("JIC_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n) continue;
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) { api.error = "libnccl.so.2 not found (set JIC_NCCL_LIB)"; return; }
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
    api.AllReduce = (decltype(api.AllReduce))dlsym(api.handle, "ncclAllReduce");
    api.AllGather = (decltype(api.AllGather))dlsym(api.handle, "ncclAllGather");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.AllGather || !api.CommDestroy) api.error = "libnccl is missing symbols";
  });
  return api;
}

template <typename R>
struct EngineT : Engine {
  jic_params prm;
  std::vector<jic_species> species;
  DevParams<R> dp;
  int device = 0, n_sm = 148;
  bool initialized = false;
  // particles (SoA)
  R *xh = nullptr, *yh = nullptr, *zh = nullptr, *vx = nullptr, *vy = nullptr, *vz = nullptr, *v_init = nullptr;
  // grid
  R *acc = nullptr, *acc2 = nullptr, *F = nullptr;   // raw deposit grid (two of them when the multi-CTA field kernel is on):
                                                     // (G,4) [Jx,Jy,Jz,rho] followed by (G) rho on the faces (field_solver != 0)
  double *gauss_h = nullptr, *ExC = nullptr;         // field_solver != 0: circulant kernel of the spectral solve, E_x of the step
  size_t gauss_smem = 0;
  double *E = nullptr, *B = nullptr, *E_int = nullptr, *B_int = nullptr, *J = nullptr, *rho = nullptr, *extE = nullptr, *extB = nullptr;
  double *s0 = nullptr, *s1 = nullptr, *E0 = nullptr, *B0 = nullptr;
  double *E2 = nullptr, *B2 = nullptr;  // ping-pong partners of E, B (multi-CTA field kernel)
  unsigned* mc_done = nullptr;
  bool mc = false;                      // multi-CTA field kernel in use
  int mc_S = 0, mc_H = 0, mc_NC = 0;
  size_t mc_smem = 0;
  int par = 0;                          // which of the ping-pong buffers the next step reads (mc only)
  RunControl* ctl = nullptr;
  // multi-GPU
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  // fused reduction over peer memory: acc, acc2, the peer flags and the step counter live in ONE allocation (`shared_block`)
  // that every other rank maps through CUDA IPC
  unsigned char* shared_block = nullptr;
  size_t acc_bytes = 0;                    // bytes of one raw grid (G x 5 reals, rounded up to 256)
  unsigned* peer_flags_local = nullptr;
  unsigned long long* fused_seq = nullptr;
  int* dev_error = nullptr;
  void* peer_block[JIC_MAX_PEERS] = {};    // mapped shared_block of every rank (own entry = shared_block)
  bool p2p = false;
  int* barrier_word = nullptr;
  // captured time loops, keyed by the number of steps (the kernels read the output pointers from RunControl at run time)
  std::map<int, cudaGraphExec_t> graphs;
  bool shared_grid = false;
  size_t shared_bytes = 0;
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int field_smem_comps = 0;
  size_t field_smem_bytes = 0;
  BinnedStore<R> bins;  // BINNED engine state (unused for INDEXED)
  // implicit Crank-Nicolson stepper (time_evolution_algorithm = 1)
  bool cn = false;
  CnState<R> cn_s[2] = {};          // double-buffered particle state, swapped every step (`par`)
  R* cn_stag = nullptr;             // staggered positions of the previous Picard iteration, (substeps, N)
  double *cn_Eg = nullptr, *cn_Bnext = nullptr, *cn_Eavg = nullptr, *cn_Bavg = nullptr;
  CnControl* cn_ctl = nullptr;
  uint8_t* cn_alive = nullptr;      // 0 = absorbed by the start-up half step (charge 0 for the whole run)
  // cell-sorted Crank-Nicolson push (csrc/jic_cn_sorted.cuh): permutation (slot -> input index), species and alive bytes per slot,
  // their scatter targets, and the counting sort's histogram / offsets
  bool cn_sorted = false;
  int cn_sort_every = 4, cn_age = 0;  // sort at steps 0, every, 2 every ... since initialisation (JIC_CN_SORT_EVERY)
  int *cn_perm = nullptr, *cn_perm2 = nullptr;
  uint8_t *cn_sp = nullptr, *cn_sp2 = nullptr, *cn_alive2 = nullptr;
  unsigned *cn_hist = nullptr, *cn_off = nullptr;
  double* cn_EB = nullptr;  // (G,6) packed gather table of the sorted push: CnFieldArgs::EB

  int dtype() const override { return prm.dtype; }
  long long n_particles() const override { return dp.N; }
  int n_grid() const override { return dp.G; }

  template <typename T>
  int alloc(T** p, size_t n) {
    JIC_CUDA(DevicePool::get().malloc((void**)p, (n ? n : 1) * sizeof(T)));
    JIC_CUDA(cudaMemset(*p, 0, (n ? n : 1) * sizeof(T)));
    return JIC_OK;
  }

  int create(const jic_params& params, const jic_species* sp) {
    prm = params;
    if (prm.device >= 0) JIC_CUDA(cudaSetDevice(prm.device));
    JIC_CUDA(cudaGetDevice(&device));
    JIC_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
    species.assign(sp, sp + prm.n_species);
    memset(&dp, 0, sizeof(dp));
    long long n = 0;
    for (int s = 0; s < prm.n_species; ++s) {
      if (sp[s].count < 0) return fail(JIC_ERR_INVALID_ARGUMENT, "negative species count");
      n += sp[s].count;
      dp.sp_end[s] = n;
      dp.sp_q[s] = (R)sp[s].charge;
      dp.sp_m[s] = (R)sp[s].mass;
      dp.sp_qm[s] = (R)sp[s].charge_to_mass;
    }
    dp.N = n;
    dp.G = prm.n_grid;
    dp.n_species = prm.n_species;
    dp.pbl = prm.particle_bc_left; dp.pbr = prm.particle_bc_right; dp.fbl = prm.field_bc_left; dp.fbr = prm.field_bc_right;
    dp.relativistic = prm.relativistic;
    dp.track_yz = prm.track_yz;
    dp.stag = (prm.field_solver != 0 && prm.time_evolution_algorithm != 1) ? 1 : 0;  // (CN_step has no field_solver branch)
    const double Ly = prm.length_y > 0 ? prm.length_y : prm.length, Lz = prm.length_z > 0 ? prm.length_z : prm.length;
    dp.L = (R)prm.length; dp.Ly = (R)Ly; dp.Lz = (R)Lz;
    dp.half_L = (R)(prm.length / 2); dp.half_Ly = (R)(Ly / 2); dp.half_Lz = (R)(Lz / 2);
    dp.dx = (R)prm.dx; dp.inv_dx = (R)(1.0 / prm.dx); dp.half_dx = (R)(prm.dx / 2);
    dp.dt = (R)prm.dt; dp.half_dt = (R)(prm.dt / 2);
    dp.g0 = (R)prm.grid_first; dp.gl = (R)prm.grid_last;
    dp.gs = (R)(prm.grid_first - prm.dx / 2);
    dp.park_left = (R)(prm.grid_first - 1.5 * prm.dx);
    dp.park_left_cell = reference_floor_div((prm.grid_first - 1.5 * prm.dx) - (prm.grid_first - prm.dx / 2), prm.dx);
    dp.park_right = (R)(prm.grid_last + 3 * prm.dx);
    const size_t N = (size_t)n, G = (size_t)prm.n_grid;
    int rc;
    cn = prm.time_evolution_algorithm == 1;
    if (cn) {
      for (int k = 0; k < 2; ++k)
        if ((rc = alloc(&cn_s[k].x, N)) || (rc = alloc(&cn_s[k].y, N)) || (rc = alloc(&cn_s[k].z, N)) || (rc = alloc(&cn_s[k].vx, N)) ||
            (rc = alloc(&cn_s[k].vy, N)) || (rc = alloc(&cn_s[k].vz, N)))
          return rc;
      if ((rc = alloc(&cn_stag, N * (size_t)prm.cn_substeps)) || (rc = alloc(&v_init, 3 * N)) || (rc = alloc(&cn_ctl, 1)) || (rc = alloc(&cn_alive, N))) return rc;
      if ((rc = alloc(&cn_Eg, G * 3)) || (rc = alloc(&cn_Bnext, G * 3)) || (rc = alloc(&cn_Eavg, G * 3)) || (rc = alloc(&cn_Bavg, G * 3))) return rc;
      {
        // large runs: sorted push with warp-aggregated deposit (JIC_CN_SORTED_MIN overrides the particle-count threshold; 0 = always)
        long long min_n = 200000;
        if (const char* env = getenv("JIC_CN_SORTED_MIN")) min_n = atoll(env);
        if (const char* env = getenv("JIC_CN_SORT_EVERY")) cn_sort_every = atoi(env) < 1 ? 1 : (atoi(env) > 64 ? 64 : atoi(env));
        cn_sorted = (long long)N >= min_n && G >= 8 && N < (1ull << 31);
        if (cn_sorted) {
          if ((rc = alloc(&cn_perm, N)) || (rc = alloc(&cn_perm2, N)) || (rc = alloc(&cn_sp, N)) || (rc = alloc(&cn_sp2, N)) ||
              (rc = alloc(&cn_alive2, N)) || (rc = alloc(&cn_hist, G)) || (rc = alloc(&cn_off, G + 1)) || (rc = alloc(&cn_EB, G * 6)))
            return rc;
          JIC_CUDA(cudaFuncSetAttribute(k_cn_push_sorted<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cn_sorted_smem_bytes<R>()));
        }
      }
    } else if (prm.engine == JIC_ENGINE_INDEXED) {
      if ((rc = alloc(&xh, N)) || (rc = alloc(&vx, N)) || (rc = alloc(&vy, N)) || (rc = alloc(&vz, N))) return rc;
      if (prm.track_yz && ((rc = alloc(&yh, N)) || (rc = alloc(&zh, N)))) return rc;
      if ((rc = alloc(&v_init, 3 * N))) return rc;
    } else {
      if ((rc = bins.create(*this, dp, prm, n_sm))) return rc;
    }
    {
      acc_bytes = (G * (kAccRow + 1) * sizeof(R) + 255) & ~(size_t)255;
      if ((rc = alloc(&shared_block, 2 * acc_bytes + 512))) return rc;
      acc = (R*)shared_block; acc2 = (R*)(shared_block + acc_bytes);
      peer_flags_local = (unsigned*)(shared_block + 2 * acc_bytes);
      fused_seq = (unsigned long long*)(shared_block + 2 * acc_bytes + 256);
      dev_error = (int*)(shared_block + 2 * acc_bytes + 256 + 8);
      barrier_word = (int*)(shared_block + 2 * acc_bytes + 256 + 16);
    }
    if ((rc = alloc(&F, (G + 3) * kFieldRow))) return rc;
    if ((rc = alloc(&E, G * 3)) || (rc = alloc(&B, G * 3)) || (rc = alloc(&E_int, G * 3)) || (rc = alloc(&B_int, G * 3))) return rc;
    if ((rc = alloc(&J, G * 3)) || (rc = alloc(&rho, G)) || (rc = alloc(&extE, G * 3)) || (rc = alloc(&extB, G * 3))) return rc;
    if ((rc = alloc(&s0, G * kAccRow)) || (rc = alloc(&s1, G * kAccRow)) || (rc = alloc(&E0, G * 3)) || (rc = alloc(&B0, G * 3))) return rc;
    if ((rc = alloc(&ctl, 1))) return rc;
    if ((rc = alloc(&E2, G * 3)) || (rc = alloc(&B2, G * 3)) || (rc = alloc(&mc_done, 1))) return rc;
    if (dp.stag) {
      // per-step electrostatic correction (_algorithms.py:69-78): circulant kernel of the spectral solvers, built once
      if ((rc = alloc(&gauss_h, G)) || (rc = alloc(&ExC, G))) return rc;
      gauss_smem = 2 * G * sizeof(double);
      if (gauss_smem > (size_t)max_smem_optin() - 1024) return fail(JIC_ERR_UNSUPPORTED, "field_solver != 0 needs 16 G bytes of shared memory: grid too large");
      if (gauss_smem > 48 * 1024) JIC_CUDA(cudaFuncSetAttribute(k_gauss<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gauss_smem));
      k_gauss_kernel<<<(int)((G + 127) / 128), 128>>>((int)G, prm.dx, gauss_h);
      JIC_CUDA(cudaGetLastError());
      JIC_CUDA(cudaDeviceSynchronize());
    }
    {
      // multi-CTA field kernel: slices of S nodes with a halo of H = filter reach + 2; needs >= 2 slices of >= H + 2 nodes
      long long reach = 0;
      if (prm.filter_passes > 0) {
        const int sweeps = (prm.filter_passes - 1 < 16 ? prm.filter_passes - 1 : 16) + 1;
        for (int i = 0; i < prm.n_filter_strides; ++i) reach += (long long)sweeps * prm.filter_strides[i];
      }
      const char* env = getenv("JIC_FIELDS_MC");  // "0" forces the single-CTA kernel
      mc = false;
      if (cn) env = "0";  // the CN stepper has its own single-CTA field kernel
      // (one periodic and one non-periodic field boundary couples the two ends of the domain through the curl ghosts,
      //  _boundary_conditions.py:148-207: that case stays on the single-CTA kernel)
      const bool mixed = (prm.field_bc_left == JIC_BC_PERIODIC) != (prm.field_bc_right == JIC_BC_PERIODIC);
      if (!(env && env[0] == '0') && !mixed && reach + 2 < (long long)G) {
        mc_H = (int)reach + 2;
        int nc = 16;
        while (nc >= 2) {
          const int S = (int)((G + nc - 1) / nc);
          const long long last = (long long)G - (long long)(nc - 1) * S;
          const size_t smem = ((size_t)8 * (S + 2 * mc_H) + (size_t)6 * (S + 4)) * sizeof(double);
          if (S >= 2 * mc_H && S >= 64 && last >= 1 && smem <= (size_t)max_smem_optin() - 1024) { mc = true; mc_S = S; mc_NC = nc; mc_smem = smem; break; }
          nc >>= 1;
        }
        if (mc && mc_smem > 48 * 1024)
          JIC_CUDA(cudaFuncSetAttribute(k_fields_mc<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mc_smem));
      }
    }
    // deposition target of the INDEXED kernel
    shared_bytes = G * kAccRow * sizeof(R);
    int max_smem = 0;
    JIC_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    const bool fits = shared_bytes <= (size_t)max_smem;
    if (prm.deposit == JIC_DEPOSIT_SHARED_GRID && !fits) return fail(JIC_ERR_UNSUPPORTED, "grid does not fit in shared memory");
    shared_grid = prm.deposit == JIC_DEPOSIT_SHARED_GRID || (prm.deposit == JIC_DEPOSIT_AUTO && fits && N >= 4 * G);
    if (shared_grid) {
      JIC_CUDA(cudaFuncSetAttribute(k_step<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shared_bytes));
      JIC_CUDA(cudaFuncSetAttribute(k_cn_push<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shared_bytes));
    }
    JIC_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
    JIC_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    JIC_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    // field kernel: filter in shared memory, as many of the four components at a time as fit (double-buffered)
    field_smem_comps = 0;
    for (int c = 4; c >= 1; c >>= 1)
      if ((size_t)2 * c * G * sizeof(double) <= (size_t)max_smem - 2048) { field_smem_comps = c; break; }
    field_smem_bytes = (size_t)2 * field_smem_comps * G * sizeof(double);
    if (field_smem_bytes > 48 * 1024)
      JIC_CUDA(cudaFuncSetAttribute(k_fields<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)field_smem_bytes));
    return JIC_OK;
  }

  ~EngineT() override {
    // a context may be destroyed from a thread whose current device is another one (several devices in one process):
    // drain and free on the context's own device, then give the caller its device back
    int caller_device = -1;
    cudaGetDevice(&caller_device);
    if (caller_device != device) cudaSetDevice(device);
    cudaDeviceSynchronize();
    for (auto& g : graphs) cudaGraphExecDestroy(g.second);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (side) cudaStreamDestroy(side);
    if (p2p) {
      // peers may still be reading this rank's raw grid in their last field kernel: meet them before unmapping / freeing
      nccl_barrier(nullptr);
      cudaDeviceSynchronize();
      for (int r = 0; r < world; ++r) if (r != rank && peer_block[r]) cudaIpcCloseMemHandle(peer_block[r]);
    }
    cudaDeviceSynchronize();  // pooled buffers are freed stream-ordered: nothing of this context may still be running
    for (R* b : hstage) if (b) DevicePool::get().free(b);
    if (copy_stream) cudaStreamDestroy(copy_stream);
    for (int k = 0; k < 2; ++k) { if (ev_copied[k]) cudaEventDestroy(ev_copied[k]); if (ev_used[k]) cudaEventDestroy(ev_used[k]); }
    for (int k = 0; k < 2; ++k) { void* q[] = {cn_s[k].x, cn_s[k].y, cn_s[k].z, cn_s[k].vx, cn_s[k].vy, cn_s[k].vz}; for (void* p : q) if (p) DevicePool::get().free(p); }
    void* ptrs[] = {cn_EB, cn_perm, cn_perm2, cn_sp, cn_sp2, cn_alive2, cn_hist, cn_off, cn_alive, cn_stag, cn_Eg, cn_Bnext, cn_Eavg, cn_Bavg, cn_ctl, xh, yh, zh, vx, vy, vz, v_init, shared_block, F, E, B, E2, B2, E_int, B_int, J, rho, extE, extB, s0, s1, E0, B0, ctl, mc_done, gauss_h, ExC};
    for (void* p : ptrs) if (p) DevicePool::get().free(p);
    bins.destroy();
    if (comm && !comm_shared && nccl_api().CommDestroy) nccl_api().CommDestroy(comm);
    if (caller_device >= 0 && caller_device != device) cudaSetDevice(caller_device);
  }

  int nccl_barrier(cudaStream_t st) {
    if (world <= 1 || !comm) return JIC_OK;
    ncclResult_t r = nccl_api().AllReduce(barrier_word, barrier_word, 1, ncclInt, ncclSum, comm, st);
    return r == ncclSuccess ? JIC_OK : JIC_ERR_NCCL;
  }

  // Fused reduction set-up: every rank exports its shared block through CUDA IPC, the handles travel by NCCL all-gather, every
  // rank maps the others.  All ranks must agree (same host, distinct processes, every mapping succeeded); otherwise the NCCL
  // all-reduce stays.  JIC_P2P=0 forces the NCCL path.
  struct PeerInfo { cudaIpcMemHandle_t handle; unsigned long long host; long long pid; int ok; int pad; };

  int setup_p2p() {
    NcclApi& api = nccl_api();
    p2p = false;
    if (world > JIC_MAX_PEERS) return JIC_OK;
    PeerInfo mine;
    memset(&mine, 0, sizeof(mine));
    const char* env = getenv("JIC_P2P");
    mine.ok = !(env && env[0] == '0') && cudaIpcGetMemHandle(&mine.handle, shared_block) == cudaSuccess;
    cudaGetLastError();
    char host[256] = {0};
    gethostname(host, sizeof(host) - 1);
    unsigned long long h = 1469598103934665603ull;
    for (const char* c = host; *c; ++c) h = (h ^ (unsigned char)*c) * 1099511628211ull;
    mine.host = h; mine.pid = (long long)getpid();
    PeerInfo* d_all = nullptr;
    JIC_CUDA(cudaMalloc((void**)&d_all, sizeof(PeerInfo) * world));
    JIC_CUDA(cudaMemcpy(d_all + rank, &mine, sizeof(mine), cudaMemcpyHostToDevice));
    ncclResult_t r = api.AllGather(d_all + rank, d_all, sizeof(PeerInfo), ncclChar, comm, nullptr);
    std::vector<PeerInfo> all(world);
    cudaError_t ce = cudaMemcpy(all.data(), d_all, sizeof(PeerInfo) * world, cudaMemcpyDeviceToHost);  // (synchronises the null stream)
    cudaFree(d_all);
    if (r != ncclSuccess || ce != cudaSuccess) return fail(JIC_ERR_NCCL, "all-gather of the IPC handles failed");
    int ok = 1;
    for (int q = 0; q < world; ++q) {
      if (!all[q].ok || all[q].host != mine.host) ok = 0;
      for (int q2 = 0; q2 < q; ++q2) if (all[q].pid == all[q2].pid) ok = 0;  // IPC handles cannot be opened by their own process
    }
    for (int q = 0; q < world; ++q) peer_block[q] = nullptr;
    if (ok) {
      for (int q = 0; q < world && ok; ++q) {
        if (q == rank) { peer_block[q] = shared_block; continue; }
        if (cudaIpcOpenMemHandle(&peer_block[q], all[q].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { peer_block[q] = nullptr; ok = 0; cudaGetLastError(); }
      }
    }
    // agreement: the fused path is used only if EVERY rank mapped every peer
    int* d_ok = nullptr;
    JIC_CUDA(cudaMalloc((void**)&d_ok, sizeof(int)));
    JIC_CUDA(cudaMemcpy(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice));
    r = api.AllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, comm, nullptr);
    int agreed = 0;
    ce = cudaMemcpy(&agreed, d_ok, sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(d_ok);
    if (r != ncclSuccess || ce != cudaSuccess) agreed = 0;
    if (!agreed) {
      for (int q = 0; q < world; ++q) if (q != rank && peer_block[q]) { cudaIpcCloseMemHandle(peer_block[q]); peer_block[q] = nullptr; }
      return JIC_OK;
    }
    p2p = true;
    return JIC_OK;
  }

  bool fused() const { return p2p && mc && !dp.stag && !cn; }
  int comm_mode() const override { return world <= 1 ? 0 : (fused() ? 2 : 1); }

  // NCCL communicators are kept for the life of the process and shared by the contexts of one (device, rank, world): creating one
  // costs 0.3-3 s at 8 ranks, which would otherwise be paid by every Simulation.run().  All ranks take the same path as long as they
  // create their contexts in the same order (they do: one process per GPU running the same program).  JIC_COMM_CACHE=0 disables it.
  bool comm_shared = false;
  int comm_init(const void* id, int rank_, int world_) override {
    if (world_ <= 1) { rank = 0; world = 1; return JIC_OK; }
    NcclApi& api = nccl_api();
    if (!api.error.empty()) return fail(JIC_ERR_NCCL, api.error);
    JIC_CUDA(cudaSetDevice(device));
    static std::mutex mu;
    static std::map<long long, ncclComm_t> cache;
    const char* env = getenv("JIC_COMM_CACHE");
    const bool use_cache = !(env && env[0] == '0');
    const long long key = ((long long)world_ << 40) | ((long long)rank_ << 20) | (long long)device;
    {
      std::lock_guard<std::mutex> lock(mu);
      auto it = cache.find(key);
      if (use_cache && it != cache.end()) { comm = it->second; comm_shared = true; }
    }
    if (!comm) {
      ncclUniqueId uid;
      memcpy(&uid, id, sizeof(uid));
      ncclResult_t r = api.CommInitRank(&comm, world_, uid, rank_);
      if (r != ncclSuccess) return fail(JIC_ERR_NCCL, format("ncclCommInitRank: %s", api.GetErrorString ? api.GetErrorString(r) : "?"));
      if (use_cache) {
        std::lock_guard<std::mutex> lock(mu);
        cache[key] = comm;
        comm_shared = true;
      }
    }
    rank = rank_; world = world_;
    return setup_p2p();
  }

  int set_external(const float* eE, const float* eB, cudaStream_t st) override {
    const int n = dp.G * 3;
    k_f32_to_f64<R><<<(n + 255) / 256, 256, 0, st>>>(eE, extE, n);
    k_f32_to_f64<R><<<(n + 255) / 256, 256, 0, st>>>(eB, extB, n);
    launches += 2;
    JIC_CUDA(cudaGetLastError());
    return JIC_OK;
  }

  int max_smem_optin() const {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    return v;
  }

  int grid_for(long long n, int block, int per_sm) const {
    long long b = (n + block - 1) / block;
    long long cap = (long long)n_sm * per_sm;
    if (b > cap) b = cap;
    return (int)(b < 1 ? 1 : b);
  }

  R* acc_of(int p) const { return (mc && p) ? acc2 : acc; }
  double* E_of(int p) const { return (mc && p) ? E2 : E; }
  double* B_of(int p) const { return (mc && p) ? B2 : B; }

  int allreduce(cudaStream_t st, int p, bool force = false) {
    if (world <= 1) return JIC_OK;
    if (fused() && !force) return JIC_OK;  // the field kernel sums the peers' grids itself
    NcclApi& api = nccl_api();
    R* buf = acc_of(p);
    ncclResult_t r = api.AllReduce(buf, buf, (size_t)dp.G * (kAccRow + dp.stag), sizeof(R) == 8 ? ncclDouble : ncclFloat, ncclSum, comm, st);
    if (r != ncclSuccess) return fail(JIC_ERR_NCCL, format("ncclAllReduce: %s", api.GetErrorString ? api.GetErrorString(r) : "?"));
    if (!force || !fused()) launches += 1;
    return JIC_OK;
  }

  FieldArgs<R> field_args(bool init, bool record) const {
    FieldArgs<R> a;
    memset(&a, 0, sizeof(a));
    a.G = dp.G; a.fbl = dp.fbl; a.fbr = dp.fbr; a.passes = prm.filter_passes; a.n_strides = prm.n_filter_strides; a.init = init;
    for (int i = 0; i < prm.n_filter_strides; ++i) a.strides[i] = prm.filter_strides[i];
    a.alpha = prm.filter_alpha; a.dx = prm.dx; a.dt = prm.dt;
    a.acc = acc; a.E = E; a.B = B; a.E_int = E_int; a.B_int = B_int; a.J = J; a.rho = rho; a.extE = extE; a.extB = extB; a.F = F;
    a.s0 = s0; a.s1 = s1; a.E0 = E0; a.B0 = B0; a.ctl = ctl;
    a.record = record ? 1 : 0;
    a.smem_comps = field_smem_comps;
    a.ExC = (dp.stag && !init) ? ExC : nullptr;
    return a;
  }

  // HOST buffers in: the upload is cut into chunks that alternate between two device staging buffers, and the start-up kernel of a
  // chunk runs while the next chunk is on the wire -- the device never holds a full copy of x0, v0 (4.8 GB at 1e8 particles), and
  // the start-up kernels hide behind the PCIe transfer.  (CN: plain upload, then initialize.)
  R* hstage[4] = {nullptr, nullptr, nullptr, nullptr};
  long long hstage_n = 0;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_used[2] = {nullptr, nullptr};

  int initialize_host(const void* x0h, const void* v0h, cudaStream_t st) override {
    if ((!x0h || !v0h) && dp.N > 0) return fail(JIC_ERR_INVALID_ARGUMENT, "x0/v0 is null");
    if (cn || dp.N == 0) {
      R *dx = nullptr, *dv = nullptr;
      const size_t bytes = (size_t)dp.N * 3 * sizeof(R);
      JIC_CUDA(cudaMalloc((void**)&dx, bytes ? bytes : 1));
      JIC_CUDA(cudaMalloc((void**)&dv, bytes ? bytes : 1));
      cudaMemcpyAsync(dx, x0h, bytes, cudaMemcpyHostToDevice, st);
      cudaMemcpyAsync(dv, v0h, bytes, cudaMemcpyHostToDevice, st);
      int rc = initialize(dx, dv, st);
      cudaStreamSynchronize(st);
      cudaFree(dx); cudaFree(dv);
      return rc;
    }
    long long chunk = 1ll << 23;
    if (const char* env = getenv("JIC_HOST_CHUNK")) chunk = std::max(1ll, atoll(env));
    chunk = std::min(chunk, dp.N);
    if (hstage_n < chunk) {
      for (R*& b : hstage) { if (b) { cudaDeviceSynchronize(); DevicePool::get().free(b); } b = nullptr; }
      for (R*& b : hstage) JIC_CUDA(DevicePool::get().malloc((void**)&b, (size_t)chunk * 3 * sizeof(R)));
      hstage_n = chunk;
    }
    if (!copy_stream) {
      JIC_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
      for (int k = 0; k < 2; ++k) {
        JIC_CUDA(cudaEventCreateWithFlags(&ev_copied[k], cudaEventDisableTiming));
        JIC_CUDA(cudaEventCreateWithFlags(&ev_used[k], cudaEventDisableTiming));
      }
    }
    int rc = initialize_begin(st);
    if (rc) return rc;
    if (prm.engine == JIC_ENGINE_BINNED && (rc = bins.start_begin(*this, dp, st))) return rc;
    // the staging buffers may still be read by the previous call's kernels on `st`
    JIC_CUDA(cudaEventRecord(ev_used[0], st));
    JIC_CUDA(cudaEventRecord(ev_used[1], st));
    int k = 0;
    for (long long i0 = 0; i0 < dp.N; i0 += chunk, k ^= 1) {
      const long long n = std::min(chunk, dp.N - i0);
      const size_t bytes = (size_t)n * 3 * sizeof(R);
      JIC_CUDA(cudaStreamWaitEvent(copy_stream, ev_used[k], 0));
      JIC_CUDA(cudaMemcpyAsync(hstage[2 * k], (const R*)x0h + 3 * i0, bytes, cudaMemcpyHostToDevice, copy_stream));
      JIC_CUDA(cudaMemcpyAsync(hstage[2 * k + 1], (const R*)v0h + 3 * i0, bytes, cudaMemcpyHostToDevice, copy_stream));
      JIC_CUDA(cudaEventRecord(ev_copied[k], copy_stream));
      JIC_CUDA(cudaStreamWaitEvent(st, ev_copied[k], 0));
      if (prm.engine == JIC_ENGINE_INDEXED) {
        k_start<R><<<grid_for(n, 256, 8), 256, 0, st>>>(dp, hstage[2 * k], hstage[2 * k + 1], i0, n, xh, yh, zh, vx, vy, vz, v_init, acc);
        launches += 1;
      } else if ((rc = bins.start_chunk(*this, dp, hstage[2 * k], hstage[2 * k + 1], i0, n, acc, st))) {
        return rc;
      }
      face_fix(hstage[2 * k], hstage[2 * k + 1], i0, n, st);
      JIC_CUDA(cudaEventRecord(ev_used[k], st));
    }
    if (prm.engine == JIC_ENGINE_BINNED && (rc = bins.start_end(*this, dp, st))) return rc;
    JIC_CUDA(cudaGetLastError());
    return initialize_finish(st);
  }

  int initialize_begin(cudaStream_t st) {
    // fused mode: peers may still be reading this rank's raw grid in the last field kernel of a previous run
    if (p2p && nccl_barrier(st)) return fail(JIC_ERR_NCCL, "barrier before initialize failed");
    JIC_CUDA(cudaMemsetAsync(acc, 0, (size_t)dp.G * (kAccRow + 1) * sizeof(R), st));
    JIC_CUDA(cudaMemsetAsync(acc2, 0, (size_t)dp.G * (kAccRow + 1) * sizeof(R), st));
    JIC_CUDA(cudaMemsetAsync(mc_done, 0, sizeof(unsigned), st));
    JIC_CUDA(cudaMemsetAsync(ctl, 0, sizeof(RunControl), st));
    JIC_CUDA(cudaMemsetAsync(&ctl->push_t0, 0xFF, sizeof(unsigned long long), st));
    par = 0;
    cn_age = 0;
    return JIC_OK;
  }

  int initialize_finish(cudaStream_t st) {
    int rc = allreduce(st, 0, true);  // start-up always goes through NCCL
    if (rc) return rc;
    // the step-0 face correction (face_fix) is not consumed by the start-up field kernel: it stays in the raw grid for step 0, whose
    // own reduction would count the already reduced sum once per rank -- keep the total on rank 0 only
    if (world > 1 && rank != 0 && needs_face_fix()) JIC_CUDA(cudaMemsetAsync(acc + (size_t)dp.G * kAccRow, 0, (size_t)dp.G * sizeof(R), st));
    k_fields<R><<<1, 1024, field_smem_bytes, st>>>(field_args(true, false));  // init mode is always the single-CTA kernel
    launches += 1;
    if (prm.engine == JIC_ENGINE_BINNED && (rc = bins.plan(*this, dp, st))) return rc;
    JIC_CUDA(cudaGetLastError());
    initialized = true;
    return JIC_OK;
  }

  // step-0 correction of the face deposit (csrc/jic_carry.cuh: k_start_face_fix): field_solver with one periodic and one non-periodic
  // particle wall.  Every rank corrects its own particles; initialize_finish keeps the reduced sum on one rank.
  bool needs_face_fix() const {
    return dp.stag && dp.pbl != dp.pbr && (dp.pbl == JIC_BC_PERIODIC || dp.pbr == JIC_BC_PERIODIC);
  }
  void face_fix(const R* x0, const R* v0, long long i0, long long n, cudaStream_t st) {
    if (!needs_face_fix() || n <= 0) return;
    k_start_face_fix<R><<<grid_for(n, 256, 8), 256, 0, st>>>(dp, x0, v0, i0, n, acc);
    launches += 1;
  }

  int initialize(const void* x0, const void* v0, cudaStream_t st) override {
    if ((!x0 || !v0) && dp.N > 0) return fail(JIC_ERR_INVALID_ARGUMENT, "x0/v0 is null");
    int rc = initialize_begin(st);
    if (rc) return rc;
    if (cn) return initialize_cn((const R*)x0, (const R*)v0, st);
    if (prm.engine == JIC_ENGINE_INDEXED) {
      k_start<R><<<grid_for(dp.N, 256, 8), 256, 0, st>>>(dp, (const R*)x0, (const R*)v0, 0, dp.N, xh, yh, zh, vx, vy, vz, v_init, acc);
      launches += 1;
    } else if ((rc = bins.start(*this, dp, (const R*)x0, (const R*)v0, acc, st))) {
      return rc;
    }
    face_fix((const R*)x0, (const R*)v0, 0, dp.N, st);
    JIC_CUDA(cudaGetLastError());
    return initialize_finish(st);
  }

  // The reference's scan carry instead of (x0, v0): csrc/jic_carry.cuh.  Same sequence as initialize(): zero the raw grid, one particle
  // kernel that deposits the carry's current, the reduction over ranks, k_fields in init mode (filter) -- then the carry's own E, B.
  // Crank-Nicolson carry (E, B, x_n, v_n) + which particles carry q = 0
  int load_carry_cn(const void* E_in, const void* B_in, const void* x_n, const void* v_n, const uint8_t* alive_in, cudaStream_t st) override {
    if (!cn) return fail(JIC_ERR_UNSUPPORTED, "jic_load_carry_cn needs time_evolution_algorithm = 1; the Boris carry goes through jic_load_carry");
    if (!E_in || !B_in || ((!x_n || !v_n) && dp.N > 0)) return fail(JIC_ERR_INVALID_ARGUMENT, "jic_load_carry_cn: null argument");
    int rc = initialize_begin(st);
    if (rc) return rc;
    JIC_CUDA(cudaMemsetAsync(cn_ctl, 0, sizeof(CnControl), st));
    k_cn_load<R><<<grid_for(dp.N, 256, 8), 256, 0, st>>>(dp, (const R*)x_n, (const R*)v_n, alive_in, cn_s[0], v_init, cn_alive);
    if (cn_sorted) { k_cn_meta_init<R><<<grid_for(dp.N, 256, 8), 256, 0, st>>>(dp, cn_perm, cn_sp); launches += 1; }
    const int n3 = dp.G * 3;
    k_carry_copy_fields<R><<<(n3 + 255) / 256, 256, 0, st>>>((const R*)E_in, (const R*)B_in, E, B, E0, B0, n3);
    k_cn_fields<R><<<1, 1024, 0, st>>>(cn_field_args(0, true));
    launches += 3;
    JIC_CUDA(cudaGetLastError());
    initialized = true;
    return JIC_OK;
  }

  int load_carry(const void* E_in, const void* B_in, const void* x_minus, const void* x_n, const void* x_plus, const void* v_n, cudaStream_t st) override {
    if (cn) return fail(JIC_ERR_UNSUPPORTED, "jic_load_carry: the Crank-Nicolson carry is (E, B, x, v): use jic_load_carry_cn");
    if (prm.engine != JIC_ENGINE_INDEXED) return fail(JIC_ERR_UNSUPPORTED, "jic_load_carry needs the INDEXED engine (particle order is part of the carry)");
    if (!E_in || !B_in || ((!x_minus || !x_n || !x_plus || !v_n) && dp.N > 0)) return fail(JIC_ERR_INVALID_ARGUMENT, "jic_load_carry: null argument");
    int rc = initialize_begin(st);
    if (rc) return rc;
    k_load_carry<R><<<grid_for(dp.N, 256, 8), 256, 0, st>>>(dp, (const R*)x_minus, (const R*)x_n, (const R*)x_plus, (const R*)v_n, xh, yh, zh, vx, vy, vz, v_init, acc);
    launches += 1;
    JIC_CUDA(cudaGetLastError());
    if ((rc = initialize_finish(st))) return rc;
    k_carry_fields<R><<<1, 1024, 0, st>>>(field_args(true, false), (const R*)E_in, (const R*)B_in);
    launches += 1;
    JIC_CUDA(cudaGetLastError());
    return JIC_OK;
  }

  // ---- Crank-Nicolson (csrc/jic_cn.cuh) ---------------------------------------------------------------------------
  CnFieldArgs<R> cn_field_args(int it, bool prepare_only) const {
    CnFieldArgs<R> a;
    memset(&a, 0, sizeof(a));
    a.G = dp.G; a.fbl = dp.fbl; a.fbr = dp.fbr; a.it = it; a.max_iter = prm.cn_max_iterations; a.prepare_only = prepare_only ? 1 : 0;
    a.dx = prm.dx; a.dt = prm.dt; a.tol = prm.cn_tolerance;
    a.acc = acc; a.En = E; a.Bn = B; a.Eg = cn_Eg; a.Bnext = cn_Bnext; a.Eavg = cn_Eavg; a.Bavg = cn_Bavg; a.J = J; a.rho = rho;
    a.cn = cn_ctl; a.ctl = ctl; a.EB = cn_EB;
    return a;
  }

  int initialize_cn(const R* x0, const R* v0, cudaStream_t st) {
    JIC_CUDA(cudaMemsetAsync(cn_ctl, 0, sizeof(CnControl), st));
    k_cn_start<R><<<grid_for(dp.N, 256, 8), 256, 0, st>>>(dp, x0, v0, cn_s[0], v_init, cn_alive, acc);
    launches += 1;
    if (cn_sorted) { k_cn_meta_init<R><<<grid_for(dp.N, 256, 8), 256, 0, st>>>(dp, cn_perm, cn_sp); launches += 1; }
    JIC_CUDA(cudaGetLastError());
    int rc = allreduce(st, 0, true);
    if (rc) return rc;
    // rho0 -> filtered -> E_x by the Gauss prefix sum (_state_initialization.py:371-378); the kernel then goes on to a leap-frog
    // half step that the CN carry does not want: take E^0, B^0 back from the copies it made first
    k_fields<R><<<1, 1024, field_smem_bytes, st>>>(field_args(true, false));
    JIC_CUDA(cudaMemcpyAsync(E, E0, (size_t)dp.G * 3 * sizeof(double), cudaMemcpyDeviceToDevice, st));
    JIC_CUDA(cudaMemcpyAsync(B, B0, (size_t)dp.G * 3 * sizeof(double), cudaMemcpyDeviceToDevice, st));
    k_cn_fields<R><<<1, 1024, 0, st>>>(cn_field_args(0, true));
    launches += 2;
    JIC_CUDA(cudaGetLastError());
    initialized = true;
    return JIC_OK;
  }

  // one CN step reading particle buffer `p`: max_iter x (push, all-reduce, fields); iterations after convergence return at once
  // sorted variant, a step that sorts: the state of buffer p is scattered, cell by cell, into buffer p ^ 1 (with the permutation and
  // the per-slot bytes), the Picard iterations read p ^ 1 and write p, so the step ends where it began.  A step that does not sort
  // (particles move less than a cell per step, so the order stays good for a few steps: cn_sort_every) reads p and writes p ^ 1 like
  // the unsorted stepper; slot i stays slot i, the permutation and the bytes stand.
  int enqueue_step_cn_sorted(cudaStream_t st, int p, bool sort) {
    const int g = grid_for(dp.N, 256, 8);
    const size_t N = (size_t)dp.N;
    int src = p, dst = p ^ 1;  // the iterations read src and write dst
    if (sort) {
      k_cn_hist<R><<<g, 256, 0, st>>>(dp, cn_s[p].x, cn_hist);
      k_cn_scan<<<1, 1024, 0, st>>>(dp.G, cn_hist, cn_off);
      k_cn_scatter<R><<<g, 256, 0, st>>>(dp, cn_s[p], cn_s[p ^ 1], cn_perm, cn_perm2, cn_sp, cn_sp2, cn_alive, cn_alive2, cn_off, cn_hist);
      JIC_CUDA(cudaMemsetAsync(cn_hist, 0, sizeof(unsigned) * dp.G, st));  // (the scatter used it as its cursors)
      JIC_CUDA(cudaMemcpyAsync(cn_perm, cn_perm2, N * sizeof(int), cudaMemcpyDeviceToDevice, st));
      JIC_CUDA(cudaMemcpyAsync(cn_sp, cn_sp2, N, cudaMemcpyDeviceToDevice, st));
      JIC_CUDA(cudaMemcpyAsync(cn_alive, cn_alive2, N, cudaMemcpyDeviceToDevice, st));
      launches += 3;
      src = p ^ 1; dst = p;
    }
    const int gs = grid_for(dp.N, kCnSortedThreads, JIC_CN_SORTED_MINBLOCKS);
    for (int it = 0; it < prm.cn_max_iterations; ++it) {
      k_cn_push_sorted<R><<<gs, kCnSortedThreads, cn_sorted_smem_bytes<R>(), st>>>(dp, cn_s[src], cn_s[dst], cn_stag, prm.cn_substeps, it, cn_EB,
                                                                                    acc, cn_alive, cn_sp, cn_ctl);
      int rc = allreduce(st, 0, true);
      if (rc) return rc;
      k_cn_fields<R><<<1, 1024, 0, st>>>(cn_field_args(it, false));
      launches += 2;
    }
    if (ke_hist) {
      k_cn_kinetic_sorted<R><<<g, 256, 0, st>>>(dp, cn_s[dst].vx, cn_s[dst].vy, cn_s[dst].vz, cn_sp, nullptr, ctl, 1);
      launches += 1;
    }
    k_cn_record_sorted<R><<<g, 256, 0, st>>>(dp, cn_s[dst], cn_perm, ctl);
    launches += 1;
    JIC_CUDA(cudaGetLastError());
    return JIC_OK;
  }
  // host-side cursor of the sorted stepper: (buffer that holds x_n, steps since the last sort) after one more step
  static void cn_sorted_advance(int& p, int& age, int every) {
    if (age != 0) p ^= 1;
    age = (age + 1) % every;
  }

  int enqueue_step_cn(cudaStream_t st, int p) {
    const int g = grid_for(dp.N, 256, 8);
    int per_sm = (int)((size_t)200 * 1024 / (shared_bytes + 1024));
    per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
    const int gs = grid_for(dp.N, 256, per_sm);  // persistent CTAs with a private grid each (deposit = shared)
    for (int it = 0; it < prm.cn_max_iterations; ++it) {
      if (shared_grid) k_cn_push<R, true><<<gs, 256, shared_bytes, st>>>(dp, cn_s[p], cn_s[p ^ 1], cn_stag, prm.cn_substeps, it, cn_Eavg, cn_Bavg, acc, cn_alive, cn_ctl);
      else k_cn_push<R, false><<<g, 256, 0, st>>>(dp, cn_s[p], cn_s[p ^ 1], cn_stag, prm.cn_substeps, it, cn_Eavg, cn_Bavg, acc, cn_alive, cn_ctl);
      int rc = allreduce(st, 0, true);
      if (rc) return rc;
      k_cn_fields<R><<<1, 1024, 0, st>>>(cn_field_args(it, false));
      launches += 2;
    }
    if (ke_hist) enqueue_kinetic_hist(st, p);
    k_cn_record<R><<<g, 256, 0, st>>>(dp, cn_s[p ^ 1], ctl);
    launches += 1;
    return JIC_OK;
  }

  int picard_iterations(long long* last, long long* total, cudaStream_t st) override {
    if (!cn) return fail(JIC_ERR_BAD_STATE, "not a Crank-Nicolson context");
    CnControl h;
    JIC_CUDA(cudaMemcpyAsync(&h, cn_ctl, sizeof(h), cudaMemcpyDeviceToHost, st));
    JIC_CUDA(cudaStreamSynchronize(st));
    if (last) *last = h.last_iters;
    if (total) *total = h.total_iters;
    return JIC_OK;
  }

  // enqueue one full step on `st` (used under stream capture): particle kernel(s), all-reduce, field kernel
  int enqueue_push(cudaStream_t st, int p) {
    R* acc = acc_of(p);  // (shadows the member: the raw grid this step deposits into)
    if (prm.engine == JIC_ENGINE_INDEXED) {
      if (shared_grid) {
        // persistent CTAs: one shared copy of the grid each
        const int per_sm = (int)((size_t)200 * 1024 / (shared_bytes + 1024));
        const int g = grid_for(dp.N, 256, per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm));
        k_step<R, true><<<g, 256, shared_bytes, st>>>(dp, xh, yh, zh, vx, vy, vz, F, acc, ctl);
      } else {
        k_step<R, false><<<grid_for(dp.N, 256, 8), 256, 0, st>>>(dp, xh, yh, zh, vx, vy, vz, F, acc, ctl);
      }
      launches += 1;
      return JIC_OK;
    }
    return bins.step(*this, dp, F, acc, ctl, st);
  }

  // grid part of a step.  BINNED: the plan of the next push only depends on the push that just ran, so it runs on a side
  // stream next to the all-reduce + field kernel (fork/join through events; inside a capture this becomes a parallel branch).
  FieldArgsMC<R> field_args_mc(int p) const {
    FieldArgsMC<R> a;
    memset(&a, 0, sizeof(a));
    a.G = dp.G; a.fbl = dp.fbl; a.fbr = dp.fbr; a.passes = prm.filter_passes; a.n_strides = prm.n_filter_strides;
    for (int i = 0; i < prm.n_filter_strides; ++i) a.strides[i] = prm.filter_strides[i];
    a.alpha = prm.filter_alpha; a.dx = prm.dx; a.dt = prm.dt;
    a.S = mc_S; a.H = mc_H; a.NC = mc_NC;
    a.acc_cur = acc_of(p); a.acc_next = acc_of(p ^ 1);
    a.E_r = E_of(p); a.B_r = B_of(p); a.E_w = E_of(p ^ 1); a.B_w = B_of(p ^ 1);
    a.E_int = E_int; a.B_int = B_int; a.J = J; a.rho = rho; a.extE = extE; a.extB = extB; a.F = F;
    a.record = 1; a.ctl = ctl; a.done = mc_done;
    a.ExC = dp.stag ? ExC : nullptr;
    a.world = 1; a.rank = rank; a.seq = fused_seq; a.error = dev_error;
    if (fused()) {
      a.world = world;
      const size_t flags_off = 2 * acc_bytes, acc_off = p ? acc_bytes : 0;
      for (int q = 0; q < world; ++q) {
        a.peer_acc[q] = (const R*)((unsigned char*)peer_block[q] + acc_off);
        a.peer_flags[q] = (unsigned*)((unsigned char*)peer_block[q] + flags_off);
      }
    }
    return a;
  }

  // field_solver != 0: E_x of this step from the face charge density the push just deposited (k_gauss), before the field kernel
  int enqueue_gauss(cudaStream_t st, int p) {
    GaussArgs<R> a;
    memset(&a, 0, sizeof(a));
    a.G = dp.G; a.fbl = dp.fbl; a.fbr = dp.fbr; a.passes = prm.filter_passes; a.n_strides = prm.n_filter_strides; a.mode = prm.field_solver;
    for (int i = 0; i < prm.n_filter_strides; ++i) a.strides[i] = prm.filter_strides[i];
    a.alpha = prm.filter_alpha; a.dx = prm.dx;
    a.accS = acc_of(p) + (size_t)dp.G * kAccRow; a.h = gauss_h; a.Ex = ExC;
    int nc = (dp.G + 7) / 8;  // >= 8 nodes (one per warp) per CTA
    if (nc > n_sm) nc = n_sm;
    k_gauss<R><<<nc, kGaussThreads, gauss_smem, st>>>(a);
    launches += 1;
    return JIC_OK;
  }

  int enqueue_fields(cudaStream_t st, int p) {
    const bool binned = prm.engine == JIC_ENGINE_BINNED;
    int rc;
    if (binned) {
      JIC_CUDA(cudaEventRecord(ev_fork, st));
      JIC_CUDA(cudaStreamWaitEvent(side, ev_fork, 0));
      if ((rc = bins.plan(*this, dp, side))) return rc;
      JIC_CUDA(cudaEventRecord(ev_join, side));
    }
    if ((rc = allreduce(st, p))) return rc;
    if (dp.stag && (rc = enqueue_gauss(st, p))) return rc;
    if (mc) k_fields_mc<R><<<mc_NC, kFieldsMcThreads, mc_smem, st>>>(field_args_mc(p));
    else k_fields<R><<<1, 1024, field_smem_bytes, st>>>(field_args(false, true));
    launches += 1;
    if (binned) JIC_CUDA(cudaStreamWaitEvent(st, ev_join, 0));
    return JIC_OK;
  }

  // jic_outputs.kinetic_energy: per-species kinetic energy of the velocities the push just produced (before the field kernel advances
  // the history row, and before the plan -- forked inside enqueue_fields -- flips the store's buffers)
  bool ke_hist = false;
  int enqueue_kinetic_hist(cudaStream_t st, int p) {
    if (cn) {
      k_kinetic_hist<R><<<grid_for(dp.N, 256, 8), 256, 0, st>>>(dp, cn_s[p ^ 1].vx, cn_s[p ^ 1].vy, cn_s[p ^ 1].vz, ctl, 1);
    } else if (prm.engine == JIC_ENGINE_BINNED) {
      return bins.kinetic_hist(*this, dp, ctl, st);
    } else {
      k_kinetic_hist<R><<<grid_for(dp.N, 256, 8), 256, 0, st>>>(dp, vx, vy, vz, ctl, 0);
    }
    launches += 1;
    return JIC_OK;
  }

  // one step reading the ping-pong buffers `p` (always 0 without the multi-CTA field kernel)
  int enqueue_step(cudaStream_t st, int p) {
    if (cn) return enqueue_step_cn(st, p);
    int rc = enqueue_push(st, p);
    if (rc == JIC_OK && ke_hist) rc = enqueue_kinetic_hist(st, p);
    return rc ? rc : enqueue_fields(st, p);
  }

  int begin_run(const jic_outputs& out, cudaStream_t st) {
    k_begin_run<<<1, 1, 0, st>>>(ctl, out);
    launches += 1;
    JIC_CUDA(cudaGetLastError());
    return JIC_OK;
  }

  // n real steps (no histories) with CUDA events around the particle kernel(s) and the grid part of every step
  int profile(long long n, double* ms_push, double* ms_fields, cudaStream_t st) override {
    if (!initialized) return fail(JIC_ERR_BAD_STATE, "jic_profile_steps before jic_initialize");
    jic_outputs none;
    memset(&none, 0, sizeof(none));
    ke_hist = false;
    int rc0 = begin_run(none, st);
    if (rc0) return rc0;
    std::vector<cudaEvent_t> ev(3 * (size_t)n);
    for (auto& e : ev) JIC_CUDA(cudaEventCreate(&e));
    int rc = JIC_OK;
    for (long long s = 0; s < n && rc == JIC_OK; ++s) {
      cudaEventRecord(ev[3 * s], st);
      rc = cn ? (cn_sorted ? enqueue_step_cn_sorted(st, par, cn_age == 0) : enqueue_step_cn(st, par)) : enqueue_push(st, par);
      cudaEventRecord(ev[3 * s + 1], st);
      if (rc == JIC_OK && !cn) rc = enqueue_fields(st, par);
      cudaEventRecord(ev[3 * s + 2], st);
      if (flips()) par ^= 1;
      if (cn_sorted) cn_sorted_advance(par, cn_age, cn_sort_every);
    }
    cudaError_t ce = cudaStreamSynchronize(st);
    double a = 0, b = 0;
    if (rc == JIC_OK && ce == cudaSuccess) {
      for (long long s = 0; s < n; ++s) {
        float t1 = 0, t2 = 0;
        cudaEventElapsedTime(&t1, ev[3 * s], ev[3 * s + 1]);
        cudaEventElapsedTime(&t2, ev[3 * s + 1], ev[3 * s + 2]);
        a += t1; b += t2;
      }
    }
    for (auto& e : ev) cudaEventDestroy(e);
    if (ce != cudaSuccess) return fail(JIC_ERR_CUDA, format("profile: %s", cudaGetErrorString(ce)));
    if (ms_push) *ms_push = a;
    if (ms_fields) *ms_fields = b;
    return rc;
  }

  int get_graph(int steps, cudaStream_t st, cudaGraphExec_t* exec) {
    // the buffer pointers baked into the graph depend on the starting parity (and, sorted CN, on which of its steps sort)
    const int key = ((steps * 2 + par) * 2 + (ke_hist ? 1 : 0)) * 64 + (cn_sorted ? cn_age : 0);
    auto it = graphs.find(key);
    if (it != graphs.end()) { *exec = it->second; return JIC_OK; }
    cudaStream_t cs;
    JIC_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    const long long before = launches;
    cudaError_t e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
    int rc = JIC_OK;
    if (e == cudaSuccess) {
      if (cn_sorted) {
        int p = par, age = cn_age;
        for (int s = 0; s < steps && rc == JIC_OK; ++s) {
          rc = enqueue_step_cn_sorted(cs, p, age == 0);
          cn_sorted_advance(p, age, cn_sort_every);
        }
      } else {
        for (int s = 0; s < steps && rc == JIC_OK; ++s) rc = enqueue_step(cs, flips() ? (par ^ (s & 1)) : (cn ? par : 0));
      }
    }
    cudaGraph_t graph = nullptr;
    cudaError_t e2 = cudaStreamEndCapture(cs, &graph);
    launches = before;  // capture does not launch
    cudaStreamDestroy(cs);
    if (e != cudaSuccess) return fail(JIC_ERR_CUDA, format("begin capture: %s", cudaGetErrorString(e)));
    if (rc != JIC_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e2 != cudaSuccess) return fail(JIC_ERR_CUDA, format("end capture: %s", cudaGetErrorString(e2)));
    cudaGraphExec_t ex;
    e = cudaGraphInstantiate(&ex, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail(JIC_ERR_CUDA, format("graph instantiate: %s", cudaGetErrorString(e)));
    graphs[key] = ex;
    launches_per_step = 0;
    *exec = ex;
    (void)st;
    return JIC_OK;
  }
  long long launches_per_step = 0;
  long long steps_run = 0;

  int run(long long n, const jic_outputs* outp, cudaStream_t st) override {
    if (!initialized) return fail(JIC_ERR_BAD_STATE, "jic_run before jic_initialize");
    if (n <= 0) return JIC_OK;
    jic_outputs out;
    memset(&out, 0, sizeof(out));
    if (outp) out = *outp;
    if (!cn && prm.engine == JIC_ENGINE_BINNED && (out.positions || out.velocities))
      return fail(JIC_ERR_UNSUPPORTED, "particle histories need the INDEXED engine");
    if (!cn && out.positions && !prm.track_yz) return fail(JIC_ERR_INVALID_ARGUMENT, "positions history needs track_yz=1");
    if (steps_run > 0) {
      // the kernels report trouble through sticky device flags: look at them before queueing more work
      int rc = check_status(st);
      if (rc) return rc;
    }
    steps_run += n;
    ke_hist = out.kinetic_energy != nullptr;
    if (ke_hist) JIC_CUDA(cudaMemsetAsync(out.kinetic_energy, 0, (size_t)n * dp.n_species * sizeof(double), st));
    int rcb = begin_run(out, st);
    if (rcb) return rcb;
    const int chunk = prm.steps_per_graph > 0 ? prm.steps_per_graph : 16;
    const long long per_step = count_launches_per_step();
    long long done = 0;
    while (done < n) {
      const int steps = (int)((n - done) >= chunk ? chunk : 1);
      cudaGraphExec_t ex;
      int rc = get_graph(steps, st, &ex);
      if (rc) return rc;
      JIC_CUDA(cudaGraphLaunch(ex, st));
      launches += per_step * steps;
      done += steps;
      if (flips()) par ^= steps & 1;
      if (cn_sorted) {
        for (int k = 0; k < steps; ++k) {
          if (cn_age == 0) launches += 3;  // (hist, scan, scatter of a step that sorts)
          cn_sorted_advance(par, cn_age, cn_sort_every);
        }
      }
    }
    return JIC_OK;
  }
  // do consecutive steps alternate between the ping-pong buffers?  (multi-CTA field kernel; unsorted CN state.  The sorted CN stepper
  // keeps its own cursor: cn_sorted_advance.)
  bool flips() const { return mc || (cn && !cn_sorted); }

  // device-timed push kernel: summed %globaltimer span (first CTA in, last CTA out) of the launches since the last reset
  int push_kernel_time(double* ms_sum, long long* n_launches, int reset, cudaStream_t st) override {
    RunControl h;
    JIC_CUDA(cudaMemcpyAsync(&h, ctl, sizeof(h), cudaMemcpyDeviceToHost, st));
    JIC_CUDA(cudaStreamSynchronize(st));
    if (ms_sum) *ms_sum = (double)h.push_ns * 1e-6;
    if (n_launches) *n_launches = (long long)h.push_launches;
    if (reset) {
      JIC_CUDA(cudaMemsetAsync(&ctl->push_ns, 0, 2 * sizeof(unsigned long long), st));
    }
    return JIC_OK;
  }

  int store_stats(long long out[8], cudaStream_t st) override {
    if (cn) {  // Crank-Nicolson contexts keep no binned store: out[0] says which push they run
      for (int i = 0; i < 8; ++i) out[i] = 0;
      out[0] = cn_sorted ? 1 : 0;
      return JIC_OK;
    }
    if (prm.engine != JIC_ENGINE_BINNED) return fail(JIC_ERR_UNSUPPORTED, "jic_store_stats needs the BINNED engine");
    PlanHeader h;
    JIC_CUDA(cudaMemcpyAsync(&h, bins.bd.hdr, sizeof(h), cudaMemcpyDeviceToHost, st));
    JIC_CUDA(cudaStreamSynchronize(st));
    out[0] = h.n_items; out[1] = h.ov_n[0]; out[2] = h.ov_n[1]; out[3] = h.error; out[4] = h.gen_last; out[5] = h.n_stored;
    out[6] = bins.bd.cap_total; out[7] = h.n_absorbed;
    return JIC_OK;
  }

  // sticky device-side error flags (synchronises the stream): a peer that missed the fused barrier, exhausted store capacity
  int check_status(cudaStream_t st) override {
    if (fused()) {
      int derr = 0;
      JIC_CUDA(cudaMemcpyAsync(&derr, dev_error, sizeof(int), cudaMemcpyDeviceToHost, st));
      JIC_CUDA(cudaStreamSynchronize(st));
      if (derr == 3) return fail(JIC_ERR_BAD_STATE, "fused reduction: a peer rank did not reach the step within the spin limit");
    }
    if (!cn && prm.engine == JIC_ENGINE_BINNED && initialized) return bins.check_error(*this, st);
    JIC_CUDA(cudaStreamSynchronize(st));
    return JIC_OK;
  }

  long long count_launches_per_step() const {
    if (cn) return 1 + (ke_hist ? 1 : 0) + (long long)prm.cn_max_iterations * (2 + (world > 1 ? 1 : 0));
    long long k = 2 + ((world > 1 && !fused()) ? 1 : 0) + dp.stag + (ke_hist ? 1 : 0);
    if (prm.engine == JIC_ENGINE_BINNED) k += bins.extra_launches_per_step();
    return k;
  }

  int get_fields(void* Eo, void* Bo, void* Jo, void* rhoo, cudaStream_t st) override {
    const long long n3 = (long long)dp.G * 3;
    if (Eo) k_convert<double, R><<<(int)((n3 + 255) / 256), 256, 0, st>>>(cn ? E : E_int, (R*)Eo, n3);  // (CN keeps E, B at integer time)
    if (Bo) k_convert<double, R><<<(int)((n3 + 255) / 256), 256, 0, st>>>(cn ? B : B_int, (R*)Bo, n3);
    if (Jo) k_convert<double, R><<<(int)((n3 + 255) / 256), 256, 0, st>>>(J, (R*)Jo, n3);
    if (rhoo) k_convert<double, R><<<(dp.G + 255) / 256, 256, 0, st>>>(rho, (R*)rhoo, dp.G);
    JIC_CUDA(cudaGetLastError());
    return JIC_OK;
  }

  int get_initial(void* E0o, void* B0o, void* vinit, cudaStream_t st) override {
    const long long n3 = (long long)dp.G * 3;
    if (E0o) k_convert<double, R><<<(int)((n3 + 255) / 256), 256, 0, st>>>(E0, (R*)E0o, n3);
    if (B0o) k_convert<double, R><<<(int)((n3 + 255) / 256), 256, 0, st>>>(B0, (R*)B0o, n3);
    if (vinit) {
      if (!cn && prm.engine != JIC_ENGINE_INDEXED) return fail(JIC_ERR_UNSUPPORTED, "initial_velocities need the INDEXED engine");
      JIC_CUDA(cudaMemcpyAsync(vinit, v_init, (size_t)dp.N * 3 * sizeof(R), cudaMemcpyDeviceToDevice, st));
    }
    JIC_CUDA(cudaGetLastError());
    return JIC_OK;
  }

  int get_particles(void* x, void* v, uint8_t* alive, cudaStream_t st) override {
    if (cn && cn_sorted) {  // x_n, v_n back in input order through the permutation
      k_cn_export_sorted<R><<<grid_for(dp.N, 256, 8), 256, 0, st>>>(dp, cn_s[par], cn_perm, (R*)x, (R*)v, alive);
      JIC_CUDA(cudaGetLastError());
      return JIC_OK;
    }
    if (cn) {  // x_n, v_n in input order
      DevParams<R> d = dp;
      d.track_yz = 1;
      const CnState<R>& s = cn_s[par];
      k_export_particles<R><<<grid_for(dp.N, 256, 8), 256, 0, st>>>(d, s.x, s.y, s.z, s.vx, s.vy, s.vz, (R*)x, (R*)v, alive);
      JIC_CUDA(cudaGetLastError());
      return JIC_OK;
    }
    if (prm.engine == JIC_ENGINE_BINNED) return bins.export_particles(*this, dp, (R*)x, (R*)v, alive, st);
    k_export_particles<R><<<grid_for(dp.N, 256, 8), 256, 0, st>>>(dp, xh, yh, zh, vx, vy, vz, (R*)x, (R*)v, alive);
    JIC_CUDA(cudaGetLastError());
    return JIC_OK;
  }

  int kinetic(double* out, cudaStream_t st) override {
    JIC_CUDA(cudaMemsetAsync(out, 0, sizeof(double), st));
    if (cn && cn_sorted) {
      k_cn_kinetic_sorted<R><<<grid_for(dp.N, 256, 8), 256, 0, st>>>(dp, cn_s[par].vx, cn_s[par].vy, cn_s[par].vz, cn_sp, out, nullptr, 0);
      JIC_CUDA(cudaGetLastError());
      return JIC_OK;
    }
    if (cn) {
      k_kinetic<R><<<grid_for(dp.N, 256, 8), 256, 0, st>>>(dp, cn_s[par].vx, cn_s[par].vy, cn_s[par].vz, out);
      JIC_CUDA(cudaGetLastError());
      return JIC_OK;
    }
    if (prm.engine == JIC_ENGINE_BINNED) return bins.kinetic(*this, dp, out, st);
    k_kinetic<R><<<grid_for(dp.N, 256, 8), 256, 0, st>>>(dp, vx, vy, vz, out);
    JIC_CUDA(cudaGetLastError());
    return JIC_OK;
  }
};

static int validate(const jic_params* p, const jic_species* sp, std::string& why) {
  if (!p || !sp) { why = "null params/species"; return JIC_ERR_INVALID_ARGUMENT; }
  if (p->struct_bytes != sizeof(jic_params)) { why = format("jic_params ABI mismatch: got %u bytes, expected %zu", p->struct_bytes, sizeof(jic_params)); return JIC_ERR_INVALID_ARGUMENT; }
  if (p->dtype != JIC_F64 && p->dtype != JIC_F32) { why = "dtype must be JIC_F64 or JIC_F32"; return JIC_ERR_INVALID_ARGUMENT; }
  if (p->engine != JIC_ENGINE_INDEXED && p->engine != JIC_ENGINE_BINNED) { why = "unknown engine"; return JIC_ERR_INVALID_ARGUMENT; }
  if (p->n_grid < 3) { why = "number_grid_points must be >= 3"; return JIC_ERR_INVALID_ARGUMENT; }
  if (p->n_species < 1 || p->n_species > JIC_MAX_SPECIES) { why = "n_species out of range"; return JIC_ERR_INVALID_ARGUMENT; }
  if (!(p->length > 0) || !(p->dx > 0) || !(p->dt > 0)) { why = "length, dx, dt must be positive"; return JIC_ERR_INVALID_ARGUMENT; }
  const int bcs[4] = {p->particle_bc_left, p->particle_bc_right, p->field_bc_left, p->field_bc_right};
  for (int b : bcs) if (b < 0 || b > 2) { why = "boundary codes are 0 (periodic), 1 (reflective), 2 (absorbing)"; return JIC_ERR_INVALID_ARGUMENT; }
  if (p->filter_passes < 0 || p->n_filter_strides < 0 || p->n_filter_strides > JIC_MAX_STRIDES) { why = "bad filter parameters"; return JIC_ERR_INVALID_ARGUMENT; }
  for (int i = 0; i < p->n_filter_strides; ++i) if (p->filter_strides[i] <= 0) { why = "filter strides must be positive"; return JIC_ERR_INVALID_ARGUMENT; }
  if (p->field_solver < 0 || p->field_solver > 3) { why = "field_solver must be 0 (none), 1 (Gauss FFT), 2 (Gauss Cartesian) or 3 (Poisson FFT)"; return JIC_ERR_INVALID_ARGUMENT; }
  if (p->time_evolution_algorithm != 0 && p->time_evolution_algorithm != 1) { why = "time_evolution_algorithm must be 0 (Boris) or 1 (Crank-Nicolson)"; return JIC_ERR_INVALID_ARGUMENT; }
  if (p->time_evolution_algorithm == 1 && (p->cn_substeps < 1 || p->cn_max_iterations < 1 || !(p->cn_tolerance >= 0))) {
    why = "Crank-Nicolson needs cn_substeps >= 1, cn_max_iterations >= 1, cn_tolerance >= 0"; return JIC_ERR_INVALID_ARGUMENT; }
  for (int i = 0; i < 2; ++i) if (p->reserved[i]) { why = "reserved fields must be zero"; return JIC_ERR_INVALID_ARGUMENT; }
  return JIC_OK;
}

}  // namespace jic

using namespace jic;

struct jic_context {
  Engine* eng;
};

extern "C" {

int jic_abi_version(void) { return JIC_ABI_VERSION; }

const char* jic_last_error(const jic_context* ctx) { return ctx && ctx->eng ? ctx->eng->error.c_str() : g_last_error.c_str(); }

int jic_create(const jic_params* params, const jic_species* species, jic_context** out) {
  if (!out) { g_last_error = "out is null"; return JIC_ERR_INVALID_ARGUMENT; }
  *out = nullptr;
  std::string why;
  int rc = validate(params, species, why);
  if (rc) { g_last_error = why; return rc; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_last_error = "no CUDA device: libjic_b200 has no CPU path"; return JIC_ERR_CUDA; }
  Engine* eng = nullptr;
  if (params->dtype == JIC_F64) { auto* e = new EngineT<double>(); rc = e->create(*params, species); eng = e; }
  else { auto* e = new EngineT<float>(); rc = e->create(*params, species); eng = e; }
  if (rc) { g_last_error = eng->error; delete eng; return rc; }
  *out = new jic_context{eng};
  return JIC_OK;
}

int jic_destroy(jic_context* ctx) {
  if (!ctx) return JIC_OK;
  delete ctx->eng;  // synchronises the context's device first
  delete ctx;
  return JIC_OK;
}

int jic_comm_unique_id(void* id) {
  NcclApi& api = nccl_api();
  if (!api.error.empty()) { g_last_error = api.error; return JIC_ERR_NCCL; }
  if (!id) { g_last_error = "jic_comm_unique_id: null buffer"; return JIC_ERR_INVALID_ARGUMENT; }
  ncclUniqueId uid;
  if (api.GetUniqueId(&uid) != ncclSuccess) { g_last_error = "ncclGetUniqueId failed"; return JIC_ERR_NCCL; }
  memcpy(id, &uid, sizeof(uid));
  return JIC_OK;
}

#define CTX_OR_FAIL(ctx) if (!(ctx) || !(ctx)->eng) { g_last_error = "null context"; return JIC_ERR_INVALID_ARGUMENT; }

int jic_comm_init(jic_context* ctx, const void* id, int rank, int world) { CTX_OR_FAIL(ctx); return ctx->eng->comm_init(id, rank, world); }
int jic_set_external_fields(jic_context* ctx, const float* eE, const float* eB, void* st) { CTX_OR_FAIL(ctx); return ctx->eng->set_external(eE, eB, (cudaStream_t)st); }
int jic_initialize(jic_context* ctx, const void* x0, const void* v0, void* st) { CTX_OR_FAIL(ctx); return ctx->eng->initialize(x0, v0, (cudaStream_t)st); }
int jic_initialize_host(jic_context* ctx, const void* x0, const void* v0, void* st) { CTX_OR_FAIL(ctx); return ctx->eng->initialize_host(x0, v0, (cudaStream_t)st); }
int jic_load_carry_cn(jic_context* ctx, const void* E, const void* B, const void* x_n, const void* v_n, const uint8_t* alive, void* st) {
  CTX_OR_FAIL(ctx);
  return ctx->eng->load_carry_cn(E, B, x_n, v_n, alive, (cudaStream_t)st);
}
int jic_load_carry(jic_context* ctx, const void* E, const void* B, const void* x_minus, const void* x_n, const void* x_plus, const void* v_n, void* st) {
  CTX_OR_FAIL(ctx);
  return ctx->eng->load_carry(E, B, x_minus, x_n, x_plus, v_n, (cudaStream_t)st);
}
int jic_run(jic_context* ctx, int64_t n, const jic_outputs* out, void* st) { CTX_OR_FAIL(ctx); return ctx->eng->run(n, out, (cudaStream_t)st); }
int jic_get_fields(jic_context* ctx, void* E, void* B, void* J, void* rho, void* st) { CTX_OR_FAIL(ctx); return ctx->eng->get_fields(E, B, J, rho, (cudaStream_t)st); }
int jic_get_initial(jic_context* ctx, void* E0, void* B0, void* v, void* st) { CTX_OR_FAIL(ctx); return ctx->eng->get_initial(E0, B0, v, (cudaStream_t)st); }
int jic_get_particles(jic_context* ctx, void* x, void* v, uint8_t* alive, void* st) { CTX_OR_FAIL(ctx); return ctx->eng->get_particles(x, v, alive, (cudaStream_t)st); }
int jic_kinetic_energy(jic_context* ctx, double* out, void* st) { CTX_OR_FAIL(ctx); return ctx->eng->kinetic(out, (cudaStream_t)st); }
int jic_profile_steps(jic_context* ctx, int64_t n, double* ms_push, double* ms_fields, void* st) { CTX_OR_FAIL(ctx); return ctx->eng->profile(n, ms_push, ms_fields, (cudaStream_t)st); }
int jic_check_status(jic_context* ctx, void* st) { CTX_OR_FAIL(ctx); return ctx->eng->check_status((cudaStream_t)st); }
int jic_push_kernel_time(jic_context* ctx, double* ms_sum, int64_t* n_launches, int32_t reset, void* st) {
  CTX_OR_FAIL(ctx);
  long long n = 0;
  int rc = ctx->eng->push_kernel_time(ms_sum, &n, reset, (cudaStream_t)st);
  if (n_launches) *n_launches = n;
  return rc;
}
void jic_trim_memory(void) { DevicePool::get().trim(); }
int jic_store_stats(jic_context* ctx, int64_t out[8], void* st) {
  CTX_OR_FAIL(ctx);
  long long tmp[8] = {0};
  int rc = ctx->eng->store_stats(tmp, (cudaStream_t)st);
  for (int i = 0; i < 8 && out; ++i) out[i] = tmp[i];
  return rc;
}
int64_t jic_launch_count(const jic_context* ctx) { return ctx && ctx->eng ? ctx->eng->launches : 0; }
int jic_get_picard_iterations(jic_context* ctx, int64_t* last, int64_t* total, void* st) {
  CTX_OR_FAIL(ctx);
  long long a = 0, b = 0;
  int rc = ctx->eng->picard_iterations(&a, &b, (cudaStream_t)st);
  if (last) *last = a;
  if (total) *total = b;
  return rc;
}
int jic_comm_mode(const jic_context* ctx) { return ctx && ctx->eng ? ctx->eng->comm_mode() : 0; }

int jic_sample_particles_slice(int32_t dtype, int32_t n_species, const jic_species_sampling* sp, const int64_t* first, const int64_t* local_count,
                               const double box[3], int32_t partitionable, void* x0, void* v0, void* stream) {
  if (!sp || !box || !x0 || !v0 || n_species < 1) { g_last_error = "jic_sample_particles: null argument"; return JIC_ERR_INVALID_ARGUMENT; }
  if (dtype != JIC_F64 && dtype != JIC_F32) { g_last_error = "dtype must be JIC_F64 or JIC_F32"; return JIC_ERR_INVALID_ARGUMENT; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_last_error = "no CUDA device: libjic_b200 has no CPU path"; return JIC_ERR_CUDA; }
  cudaStream_t st = (cudaStream_t)stream;
  long long offset = 0;
  for (int s = 0; s < n_species; ++s) {
    if (sp[s].count < 0) { g_last_error = "negative species count"; return JIC_ERR_INVALID_ARGUMENT; }
    SampleArgs a;
    memset(&a, 0, sizeof(a));
    a.count = sp[s].count; a.offset = offset;
    a.first = first ? first[s] : 0;
    a.n_local = local_count ? local_count[s] : sp[s].count;
    if (a.first < 0 || a.n_local < 0 || a.first + a.n_local > a.count) { g_last_error = "jic_sample_particles_slice: slice outside the species"; return JIC_ERR_INVALID_ARGUMENT; }
    a.seed_position = sp[s].seed_position; a.seed_velocity = sp[s].seed_velocity;
    a.partitionable = partitionable ? 1 : 0;
    for (int k = 0; k < 3; ++k) {
      a.random_positions[k] = sp[s].random_positions[k]; a.plus_minus[k] = sp[s].velocity_plus_minus[k];
      a.amp[k] = sp[s].perturbation_amplitude[k];
      a.wavenumber[k] = sp[s].perturbation_wavenumber[k] * 2 * 3.141592653589793 / box[k];  // _state_initialization.py:67
      a.vth[k] = sp[s].vth_over_c[k] * kC / sqrt(2.0);                                       // :74-76
      a.drift[k] = sp[s].drift_speed[k];
      a.box[k] = box[k];
    }
    if (a.n_local > 0) {
      long long blocks = (a.n_local + 255) / 256;
      if (blocks > 148 * 16) blocks = 148 * 16;
      if (dtype == JIC_F64) k_sample_species<double><<<(int)blocks, 256, 0, st>>>(a, (double*)x0, (double*)v0);
      else k_sample_species<float><<<(int)blocks, 256, 0, st>>>(a, (float*)x0, (float*)v0);
    }
    offset += a.n_local;
  }
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) { g_last_error = format("jic_sample_particles: %s", cudaGetErrorString(ce)); return JIC_ERR_CUDA; }
  return JIC_OK;
}

int jic_sample_particles(int32_t dtype, int32_t n_species, const jic_species_sampling* sp, const double box[3], int32_t partitionable,
                         void* x0, void* v0, void* stream) {
  return jic_sample_particles_slice(dtype, n_species, sp, nullptr, nullptr, box, partitionable, x0, v0, stream);
}

int jic_simulate_host(const jic_params* params, const jic_species* species, const void* x0_host, const void* v0_host,
                      const float* eE_host, const float* eB_host, int64_t n_steps, const jic_outputs* host_out, void* E0_host,
                      void* B0_host, void* vinit_host) {
  jic_context* ctx = nullptr;
  int rc = jic_create(params, species, &ctx);
  if (rc) return rc;
  Engine* e = ctx->eng;
  const size_t rs = params->dtype == JIC_F64 ? 8 : 4;
  const size_t N = (size_t)e->n_particles(), G = (size_t)e->n_grid(), T = (size_t)(n_steps > 0 ? n_steps : 0);
  std::vector<void*> dev;
  auto dalloc = [&](size_t bytes) -> void* { void* p = nullptr; if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr; dev.push_back(p); return p; };
  auto cleanup = [&](int code) { std::string msg = e->error; for (void* p : dev) cudaFree(p); jic_destroy(ctx); if (code) g_last_error = msg; return code; };
  cudaStream_t st = nullptr;
  float *deE = nullptr, *deB = nullptr;
  if (eE_host && !(deE = (float*)dalloc(G * 3 * 4))) { e->error = "cudaMalloc failed for the external electric field"; return cleanup(JIC_ERR_CUDA); }
  if (eB_host && !(deB = (float*)dalloc(G * 3 * 4))) { e->error = "cudaMalloc failed for the external magnetic field"; return cleanup(JIC_ERR_CUDA); }
  if (deE) cudaMemcpyAsync(deE, eE_host, G * 3 * 4, cudaMemcpyHostToDevice, st);
  if (deB) cudaMemcpyAsync(deB, eB_host, G * 3 * 4, cudaMemcpyHostToDevice, st);
  if ((rc = e->set_external(deE, deB, st))) return cleanup(rc);
  if ((rc = e->initialize_host(x0_host, v0_host, st))) return cleanup(rc);  // chunked upload overlapped with the start-up kernels
  jic_outputs d;
  memset(&d, 0, sizeof(d));
  const size_t sizes[7] = {T * G * 3 * rs, T * G * 3 * rs, T * G * 3 * rs, T * G * rs, T * N * 3 * rs, T * N * 3 * rs,
                           T * (size_t)params->n_species * sizeof(double)};
  void* const* hsrc = host_out ? (void* const*)host_out : nullptr;
  void** dptr = (void**)&d;
  for (int k = 0; k < 7 && hsrc; ++k)
    if (hsrc[k]) { dptr[k] = dalloc(sizes[k]); if (!dptr[k]) { e->error = "cudaMalloc failed for a history buffer"; return cleanup(JIC_ERR_CUDA); } }
  if ((rc = e->run(n_steps, &d, st))) return cleanup(rc);
  for (int k = 0; k < 7 && hsrc; ++k)
    if (hsrc[k]) cudaMemcpyAsync(hsrc[k], dptr[k], sizes[k], cudaMemcpyDeviceToHost, st);
  if (E0_host || B0_host || vinit_host) {
    void* dE0 = E0_host ? dalloc(G * 3 * rs) : nullptr;
    void* dB0 = B0_host ? dalloc(G * 3 * rs) : nullptr;
    void* dvi = vinit_host ? dalloc(N * 3 * rs) : nullptr;
    if ((E0_host && !dE0) || (B0_host && !dB0) || (vinit_host && !dvi)) { e->error = "cudaMalloc failed for the initial-state buffers"; return cleanup(JIC_ERR_CUDA); }
    if ((rc = e->get_initial(dE0, dB0, dvi, st))) return cleanup(rc);
    if (dE0) cudaMemcpyAsync(E0_host, dE0, G * 3 * rs, cudaMemcpyDeviceToHost, st);
    if (dB0) cudaMemcpyAsync(B0_host, dB0, G * 3 * rs, cudaMemcpyDeviceToHost, st);
    if (dvi) cudaMemcpyAsync(vinit_host, dvi, N * 3 * rs, cudaMemcpyDeviceToHost, st);
  }
  cudaError_t ce = cudaStreamSynchronize(st);
  if (ce != cudaSuccess) { e->error = format("simulate_host: %s", cudaGetErrorString(ce)); return cleanup(JIC_ERR_CUDA); }
  // a run that dropped particles (store capacity) or summed a partial grid (missing peer) must not come back as a result
  if ((rc = e->check_status(st))) return cleanup(rc);
  return cleanup(JIC_OK);
}

}  // extern "C"
