// Developer microbenchmark: two CTAs per SM, one "solver" warp each running the band sweep;
// how does the choice of the solver warp (and its hardware warp slot) affect the time?
#include <cstdio>
#include <cuda_runtime.h>
#include "band_solver.cuh"
using namespace csdo;
__global__ void __launch_bounds__(96, 2) k(long long *cyc, unsigned *wid, double *sink, int Nt, int NT, int reps, int mode) {
  extern __shared__ double sm[];
  double *L6 = sm, *dinv = L6 + 36 * NT, *vec = dinv + 6 * NT, *tmp = vec + 6 * NT;
  for (int i = threadIdx.x; i < 36 * NT; i += blockDim.x) L6[i] = 0.01 * ((i * 7) % 13) / 13.0;
  for (int i = threadIdx.x; i < 6 * NT; i += blockDim.x) { dinv[i] = 1.0; vec[i] = 1.0 + i * 1e-3; tmp[i] = 0; }
  __syncthreads();
  unsigned smid, warpid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
  const int second = blockIdx.x >= gridDim.x / 2;   // second wave of CTAs = second CTA on each SM
  int solver = 0;
  if (mode == 1) solver = second ? 1 : 0;
  if (mode == 2) solver = second ? 2 : 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per = Nt / 8;
  long long t0 = clock64();
  if (warp == solver) {
    for (int r = 0; r < reps; ++r) {
      if (lane < 8) interior_solve(L6, dinv, vec, tmp, lane * per, lane * per + per, NT);
      __syncwarp();
    }
  }
  long long t1 = clock64();
  if (warp == solver && lane == 0) { cyc[blockIdx.x] = t1 - t0; wid[blockIdx.x] = warpid | (smid << 8); }
  __syncthreads();
  sink[blockIdx.x * 96 + threadIdx.x] = tmp[threadIdx.x];
}
int main() {
  const int G = 296;
  long long *c; unsigned *w; double *s; cudaMalloc(&c, G * 8); cudaMalloc(&w, G * 4); cudaMalloc(&s, G * 96 * 8);
  const int Nt = 88, NT = 96, reps = 200;
  const int smem = 100 * 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int mode = 0; mode < 3; ++mode) {
    k<<<G, 96, smem>>>(c, w, s, Nt, NT, reps, mode); cudaDeviceSynchronize();
    long long h[G]; unsigned hw[G];
    cudaMemcpy(h, c, G * 8, cudaMemcpyDeviceToHost); cudaMemcpy(hw, w, G * 4, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < G; ++i) avg += h[i];
    avg /= G;
    printf("mode %d: avg %.1f cycles per block step; block0 sm %u warpid %u, block148 sm %u warpid %u, block1 sm %u warpid %u\n", mode,
           avg / reps / (2.0 * (Nt / 8)), hw[0] >> 8, hw[0] & 255, hw[148] >> 8, hw[148] & 255, hw[1] >> 8, hw[1] & 255);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
