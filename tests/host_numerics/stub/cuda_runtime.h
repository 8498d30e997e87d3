// Stub of <cuda_runtime.h> for tests/host_numerics: lets g++ compile the DEVICE headers of ramses_b200/csrc
// (hydro_device.cuh, real64.cuh, mhd_device.cuh) as ordinary C++ so that their formulas can be compared with the oracle on a
// machine without a GPU.  Test harness only -- nothing in the product includes this file.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__ inline
using std::copysign;
using std::max;
using std::min;
using std::fabs;
using std::fmax;
using std::fmin;
using std::pow;
using std::sqrt;

// enough of the CUDA execution model for the KERNEL bodies of amr_kernels.cuh to compile as plain functions: the test driver
// runs a block as blockDim.x OS threads that share the function-local `static` (= __shared__) storage and meet at a barrier
// in __syncthreads(); warp shuffles go through a per-warp exchange buffer (below); cp.async is a plain copy (sweep_dense.cuh)
struct rgpu_stub_dim3 { unsigned x, y, z; };
static thread_local rgpu_stub_dim3 threadIdx, blockIdx, blockDim, gridDim;   // set per emulated thread by the test driver
#define __shared__ static
#define __launch_bounds__(...)
#define __restrict__
inline void (*rgpu_stub_sync_hook)() = nullptr;     // block barrier of the emulated launch (tests/host_numerics/devnum.cpp)
inline void __syncthreads() { if (rgpu_stub_sync_hook) rgpu_stub_sync_hook(); }
inline void __syncwarp(unsigned = 0xffffffffu) {}
template <class T> inline T __ldg(const T* p) { return *p; }
// warp shuffles: identity unless the emulated launch switched warp exchange on (rgpu_stub_warp_on) -- then the 32 OS threads of
// a warp (consecutive linear thread ids) publish their value, meet at the warp's barrier, read the source lane, meet again
#include <pthread.h>
inline bool rgpu_stub_warp_on = false;
inline pthread_barrier_t rgpu_stub_warp_bar[64];
inline unsigned long long rgpu_stub_warp_buf[64][32];
inline int rgpu_stub_linear_tid() { return (int)(threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z)); }
template <class T> inline T rgpu_stub_shfl(T v, int src_of_lane(int lane, int arg), int arg) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
  if (!rgpu_stub_warp_on) return v;
  const int tid = rgpu_stub_linear_tid(), w = tid >> 5, lane = tid & 31;
  unsigned long long bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  rgpu_stub_warp_buf[w][lane] = bits;
  pthread_barrier_wait(&rgpu_stub_warp_bar[w]);
  const int src = src_of_lane(lane, arg);
  T r = v;
  if (src >= 0 && src < 32) { bits = rgpu_stub_warp_buf[w][src]; std::memcpy(&r, &bits, sizeof(T)); }
  pthread_barrier_wait(&rgpu_stub_warp_bar[w]);
  return r;
}
template <class T> inline T __shfl_down_sync(unsigned, T v, int d) { return rgpu_stub_shfl<T>(v, [](int l, int a) { return l + a < 32 ? l + a : -1; }, d); }
template <class T> inline T __shfl_up_sync(unsigned, T v, int d) { return rgpu_stub_shfl<T>(v, [](int l, int a) { return l - a; }, d); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m) { return rgpu_stub_shfl<T>(v, [](int l, int a) { return l ^ a; }, m); }
template <class T> inline T __shfl_sync(unsigned, T v, int s) { return rgpu_stub_shfl<T>(v, [](int, int a) { return a & 31; }, s); }
inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
inline int __any_sync(unsigned, int p) { return p; }
inline int __all_sync(unsigned, int p) { return p; }
typedef int cudaError_t;
typedef void* cudaStream_t;
