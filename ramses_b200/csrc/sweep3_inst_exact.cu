// explicit instantiation of the round-2 3-D dense sweep (sweep_dense3.cuh), riemann = exact
#include "sweep_dense3.cuh"
namespace rgpu {
template cudaError_t launch_sweep3<RIEMANN_EXACT>(const SweepArgs&, int, cudaStream_t, int);
}
