// Developer microbenchmark: cycles of pbcr_factor_cta / pbcr_solve_cta (csrc/pbcr_solver.cuh) for one
// agent-sized system per CTA, one CTA per SM, with a per-phase breakdown (thread 0's clock).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../csdotrajectoryplanning_b200/csrc -o pbcr_bench pbcr_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ unsigned long long g_ph[16];
#define DBG_INIT() long long dbg_t_ = clock64()
#define DBG_ACC(i) do { if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&g_ph[i], (unsigned long long)(t_ - dbg_t_)); dbg_t_ = t_; } } while (0)
#include "pbcr_solver.cuh"
using namespace csdo;

template <int MAXT>
__global__ void __launch_bounds__(MAXT, 1) bench(const double *Lg, int Nt, int NT, int reps, double *out, long long *cyc) {
  extern __shared__ double smem[];
  double *L = smem, *S = L + pbcr_L_doubles(NT), *b = S + pbcr_S_doubles(NT), *tmp = b + 6 * NT, *carry = tmp + 6 * NT;
  for (int i = threadIdx.x; i < pbcr_L_doubles(NT); i += blockDim.x) L[i] = Lg[i];
  __syncthreads();
  PbcrMem m{L, S, carry, tmp, tmp + 6 * (NT / kPM)};
  long long t0 = clock64();
  pbcr_factor_cta<true>(m, Nt);
  long long tf = clock64() - t0;
  long long acc = 0;
  for (int r = 0; r < reps; ++r) {
    if ((int)threadIdx.x < Nt)
      for (int k = 0; k < 6; ++k) b[k * NT + threadIdx.x] = 1.0 + 0.01 * ((threadIdx.x * 7 + k * 3 + r) % 17);
    __syncthreads();
    t0 = clock64();
    pbcr_solve_cta<true>(m, b, tmp, Nt, NT);
    acc += clock64() - t0;
  }
  if (threadIdx.x == 0) { cyc[2 * blockIdx.x] = tf; cyc[2 * blockIdx.x + 1] = acc / reps; }
  if ((int)threadIdx.x < Nt) for (int k = 0; k < 6; ++k) out[(size_t)blockIdx.x * 6 * NT + k * NT + threadIdx.x] = b[k * NT + threadIdx.x];
}

int main(int argc, char **argv) {
  const int Nt = argc > 1 ? atoi(argv[1]) : 256, reps = argc > 2 ? atoi(argv[2]) : 200;
  const int NT = (Nt + 31) & ~31, n = 6 * Nt;
  // banded SPD matrix H = B B' + diag (bandwidth 6), stored in block records
  std::vector<double> Bm((size_t)n * 7, 0.0), Lh(pbcr_L_doubles(NT), 0.0);
  srand(1);
  auto rnd = [] { return rand() / (double)RAND_MAX * 2 - 1; };
  for (int i = 0; i < n; ++i) for (int d = 0; d <= 6; ++d) Bm[(size_t)i * 7 + d] = (i - d >= 0) ? rnd() : 0.0;  // B[i][i-d]
  auto Hij = [&](int i, int j) {  // j <= i, i - j <= 6
    double s = 0;
    for (int k = std::max(0, i - 6); k <= j; ++k) s += Bm[(size_t)i * 7 + (i - k)] * Bm[(size_t)j * 7 + (j - k)];
    return s + (i == j ? 1.0 : 0.0);
  };
  for (int t = 0; t < Nt; ++t) for (int k = 0; k < 6; ++k) {
    const int i = 6 * t + k; double *B = pbcr_blk(Lh.data(), t);
    for (int d = 1; d <= 6; ++d) B[6 * k + d - 1] = (i - d >= 0) ? Hij(i, i - d) : 0.0;
    B[36 + k] = Hij(i, i);
  }
  int dev = 0, sms = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double *dL, *dout; long long *dc;
  cudaMalloc(&dL, Lh.size() * 8); cudaMemcpy(dL, Lh.data(), Lh.size() * 8, cudaMemcpyHostToDevice);
  cudaMalloc(&dout, (size_t)sms * 6 * NT * 8); cudaMalloc(&dc, sms * 16);
  const int smem = (pbcr_L_doubles(NT) + pbcr_S_doubles(NT) + 16 * NT) * 8;
  auto kern = NT <= 128 ? bench<128> : bench<256>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  unsigned long long z[16] = {0};
  for (int pass = 0; pass < 2; ++pass) {
    cudaMemcpyToSymbol(g_ph, z, sizeof(z));
    kern<<<sms, NT <= 128 ? std::max(64, NT) : NT, smem>>>(dL, Nt, NT, reps, dout, dc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  }
  std::vector<long long> c(2 * sms); cudaMemcpy(c.data(), dc, sms * 16, cudaMemcpyDeviceToHost);
  unsigned long long ph[16]; cudaMemcpyFromSymbol(ph, g_ph, sizeof(ph));
  std::vector<double> o(6 * NT); cudaMemcpy(o.data(), dout, 6 * NT * 8, cudaMemcpyDeviceToHost);
  printf("Nt %d: factor %lld cycles, solve %lld cycles/solve (CTA 0), smem %d B, x[0] %.6f\n", Nt, c[0], c[1], smem, o[0]);
  static const char *nm[8] = {"S1 sweeps", "S1 barrier", "S2+barrier", "BCR wide bwd", "S4 sweeps", "scatter+barrier", "BCR wide fwd", "BCR narrow"};
  for (int k = 0; k < 8; ++k) printf("  %-16s %8.0f cycles/solve\n", nm[k], (double)ph[k] / ((double)sms * reps));
  return 0;
}
