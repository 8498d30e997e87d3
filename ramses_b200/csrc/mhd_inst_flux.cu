// pass 4 of the MHD sweep (1-D Riemann problems): one instantiation per solver and slope mode
#include "mhd_dense.cuh"
namespace rgpu {
template <int R1D, bool SL>
static cudaError_t go(const MhdArgs& a, cudaStream_t st) {
  const int nt = 128;
  mhd_flux_kernel<R1D, SL><<<(unsigned)((a.nc + nt - 1) / nt), nt, 0, st>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_mhd_flux(const MhdArgs& a, int r1d, bool sl, cudaStream_t st) {
#define CASE(R) case R: return sl ? go<R, true>(a, st) : go<R, false>(a, st);
  switch (r1d) {
    CASE(MHD_LLF) CASE(MHD_ROE) CASE(MHD_HLL) CASE(MHD_HLLD) CASE(MHD_UPWIND) CASE(MHD_HYDRO)
    default: return cudaErrorInvalidValue;
  }
#undef CASE
}
}  // namespace rgpu
