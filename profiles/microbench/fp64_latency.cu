// FP64 pipe micro-benchmark for B200: dependent-issue latency and throughput of DFMA / DADD / DMUL
// as a function of independent chains per thread (ILP) and warps per SM.  nvcc -arch=sm_100a -O3 fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, int OP>
__global__ void k(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-9 + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      if (OP == 0) x[i] = __fma_rn(x[i], a, b);
      else if (OP == 1) x[i] = __dadd_rn(x[i], b);
      else x[i] = __dmul_rn(x[i], a);
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (double)(t1 - t0);
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0);
}
template <int ILP, int OP>
void run(int warps_per_sm, double* d) {
  const int iters = 4096;
  k<ILP, OP><<<148, warps_per_sm * 32>>>(d, iters, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<ILP, OP><<<148, warps_per_sm * 32>>>(d, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
  double inst = (double)iters * ILP;
  printf("op=%s ILP=%d warps/SM=%2d : %.2f cycles per dependent step, %.1f thread-instr/clk/SM, %.2f Tinstr/s chip\n",
         OP == 0 ? "DFMA" : OP == 1 ? "DADD" : "DMUL", ILP, warps_per_sm, cyc / iters, inst * warps_per_sm * 32 / cyc,
         inst * warps_per_sm * 32 * 148 / (ms * 1e-3) / 1e12);
}
int main() {
  double* d; cudaMalloc(&d, 8 * 148 * 1024);
  for (int w : {1, 4, 8, 12, 16, 32}) { run<1, 0>(w, d); run<2, 0>(w, d); run<4, 0>(w, d); run<8, 0>(w, d); }
  run<1, 1>(4, d); run<1, 2>(4, d); run<4, 1>(12, d); run<4, 2>(12, d);
  return 0;
}
