// FAST arithmetic mode of the 3-D dense sweep (rgpu_params.fast = 1), riemann = exact: the same source as the strict build
// compiled with FMA contraction (-fmad=true, see the Makefile) and the short reciprocal / square root / quotient forms of
// hydro_device.cuh (RGPU_FAST), in its own namespace so that no symbol is shared with the bit-exact build.
#define RGPU_FAST 1
#define rgpu rgpu_fast
#include "sweep_dense3.cuh"
#undef rgpu
#include <cstring>
extern "C" cudaError_t rgpu_fast_launch_sweep3_exact(const void* args, int nblocks, cudaStream_t st, int variant) {
  rgpu_fast::SweepArgs a;
  std::memcpy(&a, args, sizeof a);
  return rgpu_fast::launch_sweep3<rgpu_fast::RIEMANN_EXACT>(a, nblocks, st, variant);
}
