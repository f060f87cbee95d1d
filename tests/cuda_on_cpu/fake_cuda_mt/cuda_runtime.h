// TEST INFRASTRUCTURE ONLY -- a small multi-threaded CUDA emulation for g++ (C++20): one OS thread per CUDA thread of ONE block at a
// time, __syncthreads() = a barrier over the block, __shfl_*_sync = an exchange through a per-warp slot array with per-warp barriers,
// atomics = compare-and-swap on the bit pattern.  Blocks of a launch run one after the other (kernels that wait for other blocks are out
// of scope).  The sources are compiled from a copy in which `extern __shared__` was turned into `extern` (the arrays are defined by the
// harness) so that `__shared__` can mean `static` here: one copy per block, shared by its threads.
// With it the warp-level kernels of the INDEXED / CN paths (k_gauss, k_cn_fields' block reductions, the cell-sorted CN push) run on a CPU;
// tests/test_cuda_source_on_cpu.py compares them with the golden vectors.  Logic only: no memory model, no performance.
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static
#define __grid_constant__

struct EmuDim3 { unsigned x = 0, y = 0, z = 0; };
struct EmuDim3One { unsigned x = 1, y = 1, z = 1; };
extern thread_local EmuDim3 threadIdx;
extern EmuDim3 blockIdx;
extern EmuDim3One blockDim, gridDim;

struct EmuBlock {
  std::unique_ptr<std::barrier<>> all;
  std::vector<std::unique_ptr<std::barrier<>>> warp;
  std::vector<unsigned long long> slot;
};
extern EmuBlock* emu_block;

using std::floor; using std::fmod; using std::sqrt; using std::fabs; using std::fmax; using std::fma; using std::rint;
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline double sinpi(double x) { return std::sin(3.14159265358979323846 * x); }
inline double cospi(double x) { return std::cos(3.14159265358979323846 * x); }

inline void __syncthreads() { emu_block->all->arrive_and_wait(); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline long long clock64() { return 0; }
template <class T> inline T __ldg(const T* p) { return *p; }

template <class T>
inline T emu_shuffle(T v, unsigned src_lane) {
  static_assert(sizeof(T) <= 8, "shuffles move at most 64 bits");
  const unsigned w = threadIdx.x >> 5, base = w << 5;
  unsigned long long u = 0;
  std::memcpy(&u, &v, sizeof(T));
  emu_block->slot[threadIdx.x] = u;
  emu_block->warp[w]->arrive_and_wait();
  const unsigned src = base + (src_lane & 31u);
  u = src < blockDim.x ? emu_block->slot[src] : u;
  emu_block->warp[w]->arrive_and_wait();
  T r;
  std::memcpy(&r, &u, sizeof(T));
  return r;
}
template <class T> inline T __shfl_xor_sync(unsigned, T v, int o) { return emu_shuffle(v, (threadIdx.x & 31u) ^ (unsigned)o); }
template <class T> inline T __shfl_up_sync(unsigned, T v, int o) { const unsigned l = threadIdx.x & 31u; return emu_shuffle(v, l >= (unsigned)o ? l - o : l); }

template <class T> inline T __shfl_sync(unsigned, T v, int src) { return emu_shuffle(v, (unsigned)src); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu_block->warp[threadIdx.x >> 5]->arrive_and_wait(); }
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
// lanes of the warp (all of them take part) that hold the same value / the smallest value of the warp
template <class F>
inline void emu_warp_gather(int v, F&& each) {  // every lane publishes v, then sees the values of all lanes of its warp that exist
  const unsigned w = threadIdx.x >> 5, base = w << 5;
  emu_block->slot[threadIdx.x] = (unsigned long long)(long long)v;
  emu_block->warp[w]->arrive_and_wait();
  for (unsigned l = 0; l < 32 && base + l < blockDim.x; ++l) each(l, (int)(long long)emu_block->slot[base + l]);
  emu_block->warp[w]->arrive_and_wait();
}
inline unsigned __match_any_sync(unsigned, int v) {
  unsigned m = 0;
  emu_warp_gather(v, [&](unsigned l, int u) { m |= (u == v ? 1u : 0u) << l; });
  return m;
}
inline int __reduce_min_sync(unsigned, int v) {
  int r = v;
  emu_warp_gather(v, [&](unsigned, int u) { r = min(r, u); });
  return r;
}
struct double2 { double x, y; };
struct float2 { float x, y; };

template <class T, class U>
inline T emu_atomic_add_bits(T* p, T v) {
  static_assert(sizeof(T) == sizeof(U));
  U* q = reinterpret_cast<U*>(p);
  U old = __atomic_load_n(q, __ATOMIC_RELAXED), want;
  T cur;
  do {
    std::memcpy(&cur, &old, sizeof(T));
    const T sum = cur + v;
    std::memcpy(&want, &sum, sizeof(T));
  } while (!__atomic_compare_exchange_n(q, &old, want, false, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED));
  return cur;
}
inline double atomicAdd(double* p, double v) { return emu_atomic_add_bits<double, unsigned long long>(p, v); }
inline float atomicAdd(float* p, float v) { return emu_atomic_add_bits<float, unsigned>(p, v); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <class T> inline T atomicExch(T* p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }

// kernel<<<grid, block>>>(args...)  ->  emu_launch(grid, block, [&] { kernel(args...); });
template <class F>
inline void emu_launch(unsigned grid, unsigned block, F&& f) {
  gridDim.x = grid;
  blockDim.x = block;
  for (unsigned b = 0; b < grid; ++b) {
    blockIdx.x = b;
    EmuBlock blk;
    blk.all = std::make_unique<std::barrier<>>(block);
    for (unsigned w = 0; w * 32 < block; ++w) blk.warp.push_back(std::make_unique<std::barrier<>>(std::min(32u, block - w * 32)));
    blk.slot.assign(block, 0);
    emu_block = &blk;
    std::vector<std::thread> threads;
    for (unsigned t = 0; t < block; ++t) threads.emplace_back([&f, t] { threadIdx.x = t; f(); });
    for (auto& th : threads) th.join();
  }
  emu_block = nullptr;
}
