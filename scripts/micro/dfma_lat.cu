// Developer microbenchmark: dependent-issue latency of DFMA / LDS / shuffle on sm_100a (single warp).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double *out, long long *cyc, int iters) {
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += 32) sm[i] = 1.0 + i * 1e-9;
  __syncwarp();
  double a = 1.000001, b = 1e-9, r = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) { r = fma(r, a, b); r = fma(r, a, b); r = fma(r, a, b); r = fma(r, a, b); }
  long long t1 = clock64();
  double r0 = r, r1 = r + 1, r2 = r + 2, r3 = r + 3;
  for (int i = 0; i < iters; ++i) { r0 = fma(r0, a, b); r1 = fma(r1, a, b); r2 = fma(r2, a, b); r3 = fma(r3, a, b); }
  long long t2 = clock64();
  // dependent LDS chain (pointer chasing through shared memory)
  int idx = threadIdx.x;
  __shared__ int nxt[1024];
  for (int i = threadIdx.x; i < 1024; i += 32) nxt[i] = (i * 7 + 13) & 1023;
  __syncwarp();
  long long t3 = clock64();
  for (int i = 0; i < iters; ++i) { idx = nxt[idx]; idx = nxt[idx]; idx = nxt[idx]; idx = nxt[idx]; }
  long long t4 = clock64();
  float f = threadIdx.x;
  for (int i = 0; i < iters; ++i) { f = fmaf(f, 1.0001f, 0.5f); f = fmaf(f, 1.0001f, 0.5f); f = fmaf(f, 1.0001f, 0.5f); f = fmaf(f, 1.0001f, 0.5f); }
  long long t5 = clock64();
  double s = r;
  for (int i = 0; i < iters; ++i) { s = __shfl_xor_sync(0xffffffffu, s, 1); s = __shfl_xor_sync(0xffffffffu, s, 2); s = __shfl_xor_sync(0xffffffffu, s, 4); s = __shfl_xor_sync(0xffffffffu, s, 8); }
  long long t6 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t4 - t3; cyc[3] = t5 - t4; cyc[4] = t6 - t5; }
  out[threadIdx.x] = r + r0 + r1 + r2 + r3 + idx + f + s;
}
int main() {
  double *o; long long *c; cudaMalloc(&o, 32 * 8); cudaMalloc(&c, 8 * 8);
  int iters = 10000;
  lat<<<1, 32>>>(o, c, iters); cudaDeviceSynchronize();
  long long h[8]; cudaMemcpy(h, c, 64, cudaMemcpyDeviceToHost);
  printf("DFMA dependent: %.2f cyc/op; DFMA 4 independent chains: %.2f cyc/op; LDS dependent: %.2f cyc; FFMA dependent %.2f; SHFL(double) dependent %.2f\n",
         h[0] / (4.0 * iters), h[1] / (4.0 * iters), h[2] / (4.0 * iters), h[3] / (4.0 * iters), h[4] / (4.0 * iters));
  return 0;
}
