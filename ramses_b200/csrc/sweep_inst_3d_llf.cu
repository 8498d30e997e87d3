// explicit instantiation of the 3-D dense-box sweep, riemann = llf (all slope variants and tile heights)
#include "sweep_dense.cuh"
namespace rgpu {
template cudaError_t launch_sweep_dense<3, RIEMANN_LLF>(const SweepArgs&, int, cudaStream_t, int);
template cudaError_t launch_sweep_dense_amr<3, RIEMANN_LLF>(const SweepArgs&, int, cudaStream_t);
}
