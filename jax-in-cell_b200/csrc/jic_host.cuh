// Host-side plumbing shared by the engine translation unit and the particle stores.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <string>

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
