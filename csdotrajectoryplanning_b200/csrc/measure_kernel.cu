// FP64 FMA throughput microbenchmark: the roofline denominator of the DSQP
// kernel (FP64 is not in MEASURED_PEAKS.json, SURVEY.md section 8d asks the
// build to measure it).  8 independent DFMA chains per thread, 256 threads per
// CTA, 8 CTAs per SM: enough ILP x TLP to saturate the FP64 pipe.
#include <cuda_runtime.h>

#include "csdo_dsqp.h"

namespace {

__global__ void __launch_bounds__(256) dfma_kernel(double *out, int iters, double a, double b) {
  double r0 = threadIdx.x, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6, r7 = r0 + 7;
  for (int i = 0; i < iters; ++i) {
    r0 = fma(r0, a, b); r1 = fma(r1, a, b); r2 = fma(r2, a, b); r3 = fma(r3, a, b);
    r4 = fma(r4, a, b); r5 = fma(r5, a, b); r6 = fma(r6, a, b); r7 = fma(r7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r0 + r1 + r2 + r3 + r4 + r5 + r6 + r7;
}

}  // namespace

extern "C" int csdo_measure_fp64_peak(int device, double *tflops_out) {
  if (!tflops_out) return CSDO_ERR_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) return CSDO_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return CSDO_ERR_CUDA;
  const int grid = prop.multiProcessorCount * 8, block = 256, iters = 1 << 15;
  double *buf = nullptr;
  if (cudaMalloc(&buf, sizeof(double) * (size_t)grid * block) != cudaSuccess) return CSDO_ERR_NOMEM;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    dfma_kernel<<<grid, block>>>(buf, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(buf); return CSDO_ERR_CUDA; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * 8.0 * (double)iters * grid * block;
    if (rep > 0 && ms > 0) best = fl / (ms * 1e-3) * 1e-12 > best ? fl / (ms * 1e-3) * 1e-12 : best;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(buf);
  *tflops_out = best;
  return CSDO_OK;
}
