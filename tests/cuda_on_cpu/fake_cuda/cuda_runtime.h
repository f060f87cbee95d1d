// TEST INFRASTRUCTURE ONLY -- just enough of the CUDA language for g++ to compile the INDEXED engine's kernels
// (csrc/jic_device.cuh, jic_kernels.cuh, jic_carry.cuh) as ordinary host functions, so that their arithmetic and control flow can be
// executed on a CPU by ONE emulated thread (blockDim = gridDim = 1): every `for (i = tid; i < n; i += nt)` loop then runs in full,
// __syncthreads() and atomics degenerate to nothing / plain adds.  Kernels that rely on warp shuffles are declared but must not be
// called.  This checks the source's logic against the oracle without a GPU; it says nothing about races, memory spaces or performance.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__
#define __grid_constant__

struct EmuDim3 { unsigned x = 0, y = 0, z = 0; };
extern EmuDim3 threadIdx, blockIdx;
struct EmuDim3One { unsigned x = 1, y = 1, z = 1; };
extern EmuDim3One blockDim, gridDim;

using std::floor; using std::fmod; using std::sqrt; using std::fabs; using std::fmax; using std::fma; using std::rint;
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }

inline void __syncthreads() {}
inline void __threadfence() {}
inline long long clock64() { return 0; }
template <class T> inline T atomicAdd(T* p, T v) { T old = *p; *p = old + v; return old; }
template <class T> inline T atomicExch(T* p, T v) { T old = *p; *p = v; return old; }
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __shfl_xor_sync(unsigned, T, int) { std::abort(); }  // one emulated thread has no warp
template <class T> inline T __shfl_up_sync(unsigned, T, int) { std::abort(); }
inline double sinpi(double x) { return std::sin(3.14159265358979323846 * x); }
inline double cospi(double x) { return std::cos(3.14159265358979323846 * x); }
