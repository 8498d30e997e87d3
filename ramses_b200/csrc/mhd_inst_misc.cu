// passes 1-3 and 6 of the MHD sweep, the stand-alone Courant scan and the boundary kernel
#define MHD_DEFINE_KERNELS
#include "mhd_dense.cuh"
namespace rgpu {
cudaError_t launch_mhd_flux(const MhdArgs& a, int r1d, bool sl, cudaStream_t st);
cudaError_t launch_mhd_emf(const MhdArgs& a, int r2d, bool sl, cudaStream_t st);

cudaError_t launch_mhd_sweep(const MhdArgs& a, int r1d, int r2d, bool sl, int nb_update, cudaStream_t st) {
  const int nt = 256;
  const unsigned nb = (unsigned)((a.nc + nt - 1) / nt);
  mhd_prim_kernel<<<nb, nt, 0, st>>>(a);
  mhd_efield_kernel<<<nb, nt, 0, st>>>(a);
  if (sl) mhd_trace_kernel<true><<<nb, nt, 0, st>>>(a);
  else mhd_trace_kernel<false><<<nb, nt, 0, st>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  e = launch_mhd_flux(a, r1d, sl, st);
  if (e != cudaSuccess) return e;
  e = launch_mhd_emf(a, r2d, sl, st);
  if (e != cudaSuccess) return e;
  mhd_update_kernel<<<nb_update, 256, 0, st>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_mhd_courant(const double* u, const DenseGeom& g, const MPhys& P, double dx, double* part, int nb, cudaStream_t st) {
  mhd_courant_kernel<<<nb, 256, 0, st>>>(u, g, P, dx, part);
  return cudaGetLastError();
}
cudaError_t launch_mhd_boundary(double* u, const MhdBoundArgs& b, cudaStream_t st) {
  const int nthr = b.n * 8;
  mhd_boundary_kernel<<<(nthr + 127) / 128, 128, 0, st>>>(u, b);
  return cudaGetLastError();
}
}  // namespace rgpu
