// pass 5 of the MHD sweep (2-D Riemann problems at the cell edges): one instantiation per solver and slope mode
#include <cstdlib>
#include "mhd_dense.cuh"
namespace rgpu {
// MINB: resident CTAs per SM the register allocation is bounded for (see mhd_inst_flux.cu); RGPU_MHD_MINB2 overrides.
static int minb_choice(int r2d) {
  static int env = -1;
  if (env < 0) { const char* e = getenv("RGPU_MHD_MINB2"); env = e ? atoi(e) : 0; }
  if (env >= 2 && env <= 4) return env;
  return (r2d == MHD2D_ROE) ? 4 : (r2d == MHD2D_HLLD) ? 3 : 2;
}
template <int R2D, bool SL>
static cudaError_t go(const MhdArgs& a, cudaStream_t st) {
  const int nt = 128;
  const unsigned nb = (unsigned)((a.nc + nt - 1) / nt);
  const int mb = minb_choice(R2D);
  if (mb == 4) mhd_emf_kernel<R2D, SL, 4><<<nb, nt, 0, st>>>(a);
  else if (mb == 3) mhd_emf_kernel<R2D, SL, 3><<<nb, nt, 0, st>>>(a);
  else mhd_emf_kernel<R2D, SL, 2><<<nb, nt, 0, st>>>(a);
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
