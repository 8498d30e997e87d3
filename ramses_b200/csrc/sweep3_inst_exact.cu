// explicit instantiation of the round-2 3-D dense sweeps (sweep_dense3.cuh; sweep_dense4.cuh: tuning builds), riemann = exact
#include "sweep_dense4.cuh"
namespace rgpu {
template cudaError_t launch_sweep3<RIEMANN_EXACT>(const SweepArgs&, int, cudaStream_t, int);
template cudaError_t launch_sweep4<RIEMANN_EXACT>(const SweepArgs&, int, cudaStream_t, int);
}
