// Developer microbenchmark: the whole band_solve_warp (as compiled for the library) on one CTA / all SMs.
// (the factor storage holds synthetic values: timing only; compile with -DCSDO_DEV_TIMERS for the parts)
#include <cstdio>
#include <cuda_runtime.h>
#include "band_solver.cuh"
using namespace csdo;
__global__ void __launch_bounds__(96, 2) k(long long *cyc, double *sink, int Nt, int NT, int reps, int warp_sel) {
  extern __shared__ double sm[];
  BandMem bm;
  bm.L6 = sm; bm.dinv = sm + 36 * NT + kSkewPad; 
  double *rhs = bm.dinv + 6 * NT, *tmp = rhs + 6 * NT;
  bm.Sinv = tmp + 6 * NT; bm.sv = bm.Sinv + kL2Doubles; bm.G = tmp;
  __shared__ int tab[kMaxP];
  if (threadIdx.x == 0) fill_skew_table(Nt, tab);
  bm.tab = tab;
  for (int i = threadIdx.x; i < 36 * NT + kSkewPad; i += blockDim.x) bm.L6[i] = 0.01 * ((i * 7) % 13) / 13.0;
  for (int i = threadIdx.x; i < 6 * NT; i += blockDim.x) { bm.dinv[i] = 1.0; rhs[i] = 1.0 + i * 1e-3; tmp[i] = 0; }
  for (int i = threadIdx.x; i < kL2Doubles + 3 * kMaxNs; i += blockDim.x) bm.Sinv[i] = 1e-3;
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if ((threadIdx.x >> 5) == warp_sel) band_solve_warp<true>(bm, rhs, tmp, Nt, NT);
    __syncthreads();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * 96 + threadIdx.x] = rhs[threadIdx.x];
}
int main() {
  long long *c; double *s; cudaMalloc(&c, 8 * 512); cudaMalloc(&s, 512 * 96 * 8);
  const int NT = 96, reps = 100;
  const int smem = 115360;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int grid : {1, 296})
    for (int Nt : {75, 91}) {
      k<<<grid, 96, smem>>>(c, s, Nt, NT, reps, 0); cudaDeviceSynchronize();
      long long h[512]; cudaMemcpy(h, c, 8 * grid, cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
      unsigned long long d[16]; cudaMemcpyFromSymbol(d, g_dbg, sizeof(d)); unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_dbg, z, sizeof(z));
      printf("grid %d Nt %d: %.0f cycles per solve; parts S1 %.0f S2 %.0f Sinv %.0f S3 %.0f\n", grid, Nt, avg / reps,
             (double)d[0] / reps / grid, (double)d[1] / reps / grid, (double)d[2] / reps / grid, (double)d[3] / reps / grid);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
