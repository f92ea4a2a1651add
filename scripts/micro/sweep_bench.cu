// Developer microbenchmark: cycles per block step of the band-solve sweeps (one warp; "lanes 8" = the
// partitioning make_parts() picks for the horizon, up to 16 lanes).
#include <cstdio>
#include <cuda_runtime.h>
#include "band_solver.cuh"
using namespace csdo;
__global__ void k(long long *cyc, double *sink, int Nt, int NT, int reps, int nlanes) {
  extern __shared__ double sm[];
  double *L6 = sm, *dinv = L6 + 36 * NT, *vec = dinv + 6 * NT, *tmp = vec + 6 * NT;
  for (int i = threadIdx.x; i < 36 * NT; i += blockDim.x) L6[i] = 0.01 * ((i * 7) % 13) / 13.0;
  for (int i = threadIdx.x; i < 6 * NT; i += blockDim.x) { dinv[i] = 1.0; vec[i] = 1.0 + i * 1e-3; tmp[i] = 0; }
  __shared__ int tab[kMaxP];
  if (threadIdx.x == 0) fill_skew_table(Nt, tab);
  __syncthreads();
  const int lane = threadIdx.x;
  const int per = Nt / nlanes;
  const Parts pt = make_parts(Nt, tab);
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (nlanes == 8) { if (lane < pt.P) interior_solve(L6, dinv, vec, tmp, pt.start(lane), pt.start(lane) + pt.len(lane), NT); }
    else if (lane < nlanes) interior_solve(L6, dinv, vec, tmp, lane * per, lane * per + per, NT);
    __syncwarp();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  sink[threadIdx.x] = tmp[threadIdx.x];
}
int main() {
  long long *c; double *s; cudaMalloc(&c, 64); cudaMalloc(&s, 32 * 8);
  const int NT = 96, reps = 200;
  for (int Nt : {88, 75, 91, 64}) {
  const int smem = (36 + 6 + 6 + 6) * NT * 8;
  for (int nl : {1, 8}) {
    k<<<1, 32, smem>>>(c, s, Nt, NT, reps, nl); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    const int per = Nt / nl;
    printf("Nt %d lanes %d: %.1f cycles per block step (fwd+bwd counted as 2 steps per block)\n", Nt, nl, (double)h / reps / (2.0 * (nl == 8 ? (Nt >= 80 ? (Nt - 15 + 15) / 16 : (Nt - 7 + 7) / 8) : per)));
  }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
