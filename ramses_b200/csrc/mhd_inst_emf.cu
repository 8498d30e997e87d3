// pass 5 of the MHD sweep (2-D Riemann problems at the cell edges): one instantiation per solver and slope mode
#include "mhd_dense.cuh"
namespace rgpu {
template <int R2D, bool SL>
static cudaError_t go(const MhdArgs& a, cudaStream_t st) {
  const int nt = 128;
  mhd_emf_kernel<R2D, SL><<<(unsigned)((a.nc + nt - 1) / nt), nt, 0, st>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_mhd_emf(const MhdArgs& a, int r2d, bool sl, cudaStream_t st) {
#define CASE(R) case R: return sl ? go<R, true>(a, st) : go<R, false>(a, st);
  switch (r2d) {
    CASE(MHD2D_LLF) CASE(MHD2D_ROE) CASE(MHD2D_UPWIND) CASE(MHD2D_HLL) CASE(MHD2D_HLLA) CASE(MHD2D_HLLD)
    default: return cudaErrorInvalidValue;
  }
#undef CASE
}
}  // namespace rgpu
