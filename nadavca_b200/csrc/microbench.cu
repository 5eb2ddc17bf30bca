// microbench.cu -- FP64 FMA issue-rate probe: the denominator of the ALU roofline (MEASURED_PEAKS.json only has
// HBM and bf16 tensor numbers; the DP kernels are bound by the FP64 pipe).
#include "common.cuh"
#include "kernels.h"

namespace {
__global__ void __launch_bounds__(256) fp64_fma_kernel(int iters, double *sink) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
         a7 = a0 + 7;
  const double x = 1.0000001, y = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y);
    a4 = fma(a4, x, y); a5 = fma(a5, x, y); a6 = fma(a6, x, y); a7 = fma(a7, x, y);
  }
  double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (r == 12345.678) sink[0] = r;  // keep the loop alive
}
}  // namespace

// returns elapsed milliseconds of one probe launch (8 * iters FMAs per thread, 256 threads per block)
float nvbk_fp64_fma_probe(int iters, int blocks, cudaStream_t st, double *d_sink) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, st);
  fp64_fma_kernel<<<blocks, 256, 0, st>>>(iters, d_sink);
  cudaEventRecord(e1, st);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return ms;
}
