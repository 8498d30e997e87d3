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
// in __syncthreads() -- good for kernels that use barriers only (no warp shuffles, no cp.async)
struct rgpu_stub_dim3 { unsigned x, y, z; };
static thread_local rgpu_stub_dim3 threadIdx, blockIdx, blockDim, gridDim;   // set per emulated thread by the test driver
#define __shared__ static
#define __launch_bounds__(...)
#define __restrict__
inline void (*rgpu_stub_sync_hook)() = nullptr;     // block barrier of the emulated launch (tests/host_numerics/devnum.cpp)
inline void __syncthreads() { if (rgpu_stub_sync_hook) rgpu_stub_sync_hook(); }
inline void __syncwarp(unsigned = 0xffffffffu) {}
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __shfl_down_sync(unsigned, T v, int) { return v; }
template <class T> inline T __shfl_up_sync(unsigned, T v, int) { return v; }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
template <class T> inline T __shfl_sync(unsigned, T v, int) { return v; }
inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
inline int __any_sync(unsigned, int p) { return p; }
inline int __all_sync(unsigned, int p) { return p; }
typedef int cudaError_t;
typedef void* cudaStream_t;
