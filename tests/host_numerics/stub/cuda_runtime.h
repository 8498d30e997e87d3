// Stub of <cuda_runtime.h> for tests/host_numerics: lets g++ compile the DEVICE headers of ramses_b200/csrc
// (hydro_device.cuh, real64.cuh, mhd_device.cuh) as ordinary C++ so that their formulas can be compared with the oracle on a
// machine without a GPU.  Test harness only -- nothing in the product includes this file.
#pragma once
#include <cmath>
#include <cstring>
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
using std::copysign;
using std::fabs;
using std::fmax;
using std::fmin;
using std::pow;
using std::sqrt;
