// explicit instantiation of the round-2 3-D dense sweep (sweep_dense3.cuh), riemann = hll
#include "sweep_dense3.cuh"
namespace rgpu {
template cudaError_t launch_sweep3<RIEMANN_HLL>(const SweepArgs&, int, cudaStream_t, int);
}
