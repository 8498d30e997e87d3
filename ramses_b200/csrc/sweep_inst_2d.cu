// explicit instantiations of the dense-box sweep for NDIM=2, all Riemann solvers
#include "sweep_dense.cuh"
namespace rgpu {
template cudaError_t launch_sweep_dense<2, RIEMANN_LLF>(const SweepArgs&, int, cudaStream_t, int);
template cudaError_t launch_sweep_dense<2, RIEMANN_EXACT>(const SweepArgs&, int, cudaStream_t, int);
template cudaError_t launch_sweep_dense<2, RIEMANN_ACOUSTIC>(const SweepArgs&, int, cudaStream_t, int);
template cudaError_t launch_sweep_dense<2, RIEMANN_HLLC>(const SweepArgs&, int, cudaStream_t, int);
template cudaError_t launch_sweep_dense<2, RIEMANN_HLL>(const SweepArgs&, int, cudaStream_t, int);
}
