// pass 4 of the MHD sweep (1-D Riemann problems): one instantiation per solver and slope mode
#include <cstdlib>
#include "mhd_dense.cuh"
namespace rgpu {
// MINB = resident CTAs per SM the register allocation is bounded for (2: up to 255 registers, 3: 168, 4: 128).
// Measured on B200 (profiles/r1_mhd_*): see DESIGN.md; RGPU_MHD_MINB overrides the default.
static int minb_choice(int r1d) {
  static int env = -1;
  if (env < 0) { const char* e = getenv("RGPU_MHD_MINB"); env = e ? atoi(e) : 0; }
  if (env >= 2 && env <= 6) return env;
  return r1d == MHD_ROE ? 4 : (r1d == MHD_HLLD ? 3 : 2);
}
template <int R1D, bool SL>
static cudaError_t go(const MhdArgs& a, cudaStream_t st) {
  const int nt = 128;
  const unsigned nb = (unsigned)((a.nc + nt - 1) / nt);
  const int mb = minb_choice(R1D);
  if (mb == 6 && R1D == MHD_ROE) mhd_flux_kernel<R1D, SL, (R1D == MHD_ROE ? 6 : 4)><<<nb, nt, 0, st>>>(a);
  else if (mb == 5 && R1D == MHD_ROE) mhd_flux_kernel<R1D, SL, (R1D == MHD_ROE ? 5 : 4)><<<nb, nt, 0, st>>>(a);
  else if (mb >= 4) mhd_flux_kernel<R1D, SL, 4><<<nb, nt, 0, st>>>(a);
  else if (mb == 3) mhd_flux_kernel<R1D, SL, 3><<<nb, nt, 0, st>>>(a);
  else mhd_flux_kernel<R1D, SL, 2><<<nb, nt, 0, st>>>(a);
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
