// Host-side plumbing shared by the engine translation unit and the particle stores.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <string>
#include <unordered_map>

#include "../../include/jic_b200.h"

namespace jic {

inline thread_local std::string g_last_error;

inline std::string format(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  return buf;
}

#define JIC_CUDA(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess) return fail(JIC_ERR_CUDA, format("%s -> %s", #expr, cudaGetErrorString(_e))); \
  } while (0)

// Device memory of the large buffers (particle stores, per-particle arrays).  cudaMalloc / cudaFree of 17 GB cost ~8 ms and ~20 ms per
// context at 1e8 particles -- a third of a 20-step Simulation.run().  They come from a library-owned stream-ordered memory pool per
// device instead, whose release threshold keeps freed memory in the pool: the next context of the process gets it back at once.
// jic_trim_memory() hands it back to the driver; JIC_POOL=0 switches the pool off.  Small buffers, and the block other ranks map
// through CUDA IPC, stay with cudaMalloc.
struct DevicePool {
  static constexpr size_t kMinBytes = 32u << 20;
  std::mutex mu;
  std::unordered_map<int, cudaMemPool_t> pools;      // device -> pool
  std::unordered_map<void*, int> owned;              // pointer -> device
  bool enabled() const { const char* v = getenv("JIC_POOL"); return !(v && atoi(v) == 0); }
  static DevicePool& get() { static DevicePool p; return p; }
  cudaError_t malloc(void** ptr, size_t bytes) {
    if (bytes >= kMinBytes && enabled()) {
      std::lock_guard<std::mutex> lock(mu);
      int dev = 0;
      cudaGetDevice(&dev);
      auto it = pools.find(dev);
      if (it == pools.end()) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        cudaMemPool_t pool = nullptr;
        if (cudaMemPoolCreate(&pool, &props) == cudaSuccess) {
          unsigned long long keep = ~0ull;
          cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
          it = pools.emplace(dev, pool).first;
        } else {
          (void)cudaGetLastError();
        }
      }
      if (it != pools.end()) {
        cudaError_t e = cudaMallocFromPoolAsync(ptr, bytes, it->second, cudaStreamPerThread);
        if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamPerThread);  // from here on the memory is good on every stream
        if (e == cudaSuccess) { owned[*ptr] = dev; return cudaSuccess; }
        (void)cudaGetLastError();
        cudaMemPoolTrimTo(it->second, 0);  // out of memory with a full pool: give it back and take the plain road
      }
    }
    return cudaMalloc(ptr, bytes);
  }
  // the caller has synchronised the work that used `ptr`
  void free(void* ptr) {
    if (!ptr) return;
    {
      std::lock_guard<std::mutex> lock(mu);
      auto it = owned.find(ptr);
      if (it != owned.end()) {
        owned.erase(it);
        cudaFreeAsync(ptr, cudaStreamPerThread);
        return;
      }
    }
    cudaFree(ptr);
  }
  void trim() {
    std::lock_guard<std::mutex> lock(mu);
    cudaStreamSynchronize(cudaStreamPerThread);
    for (auto& kv : pools) cudaMemPoolTrimTo(kv.second, 0);
  }
};

struct Engine {
  virtual ~Engine() {}
  std::string error;
  long long launches = 0;
  int fail(int code, const std::string& msg) { error = msg; g_last_error = msg; return code; }
  virtual int comm_init(const void* id, int rank, int world) = 0;
  virtual int set_external(const float* eE, const float* eB, cudaStream_t st) = 0;
  virtual int initialize(const void* x0, const void* v0, cudaStream_t st) = 0;
  virtual int initialize_host(const void* x0_host, const void* v0_host, cudaStream_t st) = 0;
  virtual int load_carry_cn(const void* E, const void* B, const void* x_n, const void* v_n, const uint8_t* alive, cudaStream_t st) = 0;
  virtual int load_carry(const void* E, const void* B, const void* x_minus, const void* x_n, const void* x_plus, const void* v_n, cudaStream_t st) = 0;
  virtual int run(long long n, const jic_outputs* out, cudaStream_t st) = 0;
  virtual int get_fields(void* E, void* B, void* J, void* rho, cudaStream_t st) = 0;
  virtual int get_initial(void* E0, void* B0, void* vinit, cudaStream_t st) = 0;
  virtual int get_particles(void* x, void* v, uint8_t* alive, cudaStream_t st) = 0;
  virtual int kinetic(double* out, cudaStream_t st) = 0;
  virtual int check_status(cudaStream_t st) = 0;
  virtual int store_stats(long long out[8], cudaStream_t st) = 0;
  virtual int push_kernel_time(double* ms_sum, long long* n_launches, int reset, cudaStream_t st) = 0;
  virtual int profile(long long n, double* ms_push, double* ms_fields, cudaStream_t st) = 0;
  virtual int dtype() const = 0;
  virtual long long n_particles() const = 0;
  virtual int n_grid() const = 0;
  virtual int comm_mode() const = 0;
  virtual int picard_iterations(long long* last, long long* total, cudaStream_t st) = 0;
};


}  // namespace jic
