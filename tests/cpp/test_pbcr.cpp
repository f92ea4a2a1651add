// CPU test of csrc/pbcr_solver.cuh (the CTA-wide band solver of the refine kernel): the phase functions
// are plain host/device code, a loop over the thread index stands in for the threads of a phase.  The
// result is compared with a dense Cholesky solve of the same banded SPD matrix.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "pbcr_solver.cuh"

using namespace csdo;

static void host_factor(const PbcrMem &m, int Nt, int nth) {
  const PGeom g = pbcr_geom(Nt);
  for (int tid = 0; tid < nth; ++tid)
    if (tid < g.NP) pbcr_factor_partition(m, g, tid);
  for (int tid = 0; tid < nth; ++tid)
    for (int task = tid; task < 57 * g.Ps; task += nth) pbcr_assemble_task(m, g, task / 57, task % 57);
  for (int lv = 0; lv < g.Lv; ++lv) {
    const int s = 1 << lv, nact = g.Ps >> lv, nel = (nact + 1) >> 1, nsv = nact >> 1;
    for (int tid = 0; tid < nth; ++tid)
      if (tid < nel) pbcr_bcr_eliminate(m, g, s * (2 * tid + 1), s);
    for (int tid = 0; tid < nth; ++tid)
      if (tid < nsv) pbcr_bcr_survive(m, g, s * (2 * tid + 2), s);
  }
}

static void host_solve(const PbcrMem &m, double *b, double *tmp, int Nt, int NT, int nth) {
  const PGeom g = pbcr_geom(Nt);
  if (g.Ps == 0) { pbcr_interior_solve<false>(m.L, g.part_len(0), b, b, 0, NT, nullptr); return; }
  for (int tid = 0; tid < nth; ++tid)
    if (tid < g.NP) pbcr_interior_solve<false>(m.L + tid * kGrpD, g.part_len(tid), b, tmp, kPM * tid, NT, nullptr);
  for (int tid = 0; tid < nth; ++tid)
    for (int task = tid; task < 6 * g.Ps; task += nth) pbcr_sep_rhs_task(m, g, b, tmp, NT, task / 6, task % 6);
  for (int lv = 0; lv < g.Lv; ++lv) {
    const int s = 1 << lv, nsv = (g.Ps >> lv) >> 1;
    // survivors read g of eliminated blocks only and write their own g: the order inside a level is free
    for (int tid = nth - 1; tid >= 0; --tid)
      for (int task = tid; task < 6 * nsv; task += nth) pbcr_bcr_forward_task(m, g, s, task / 6, task % 6);
  }
  for (int lv = g.Lv - 1; lv >= 0; --lv) {
    const int s = 1 << lv, nel = ((g.Ps >> lv) + 1) >> 1;
    for (int tid = 0; tid < nth; ++tid)
      for (int task = tid; task < 6 * nel; task += nth) pbcr_bcr_backward_task(m, g, s, task / 6, task % 6);
  }
  for (int tid = 0; tid < nth; ++tid)
    if (tid < g.NP) pbcr_final_partition<false>(m, g, b, NT, tid);
  for (int task = 0; task < 6 * g.Ps; ++task) b[(task % 6) * NT + g.sep_block(task / 6)] = m.xs[task];
}

int main() {
  std::mt19937_64 rng(12345);
  std::uniform_real_distribution<double> U(-1.0, 1.0);
  double worst = 0.0;
  const int sizes[] = {3, 4, 5, 6, 7, 8, 9, 11, 12, 16, 31, 33, 64, 91, 96, 127, 128, 193, 255, 256, 300, 512};
  for (int Nt : sizes) {
    const int n = 6 * Nt, NT = (Nt + 31) & ~31;
    // banded SPD matrix: H = B B' + diag, B lower banded with bandwidth 6, widely varying scales
    std::vector<double> Bm((size_t)n * n, 0.0), H((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i)
      for (int j = std::max(0, i - 6); j <= i; ++j) Bm[(size_t)i * n + j] = U(rng) * ((i % 6 == 4) ? 30.0 : 1.0);
    for (int i = 0; i < n; ++i)
      for (int j = std::max(0, i - 6); j <= i; ++j) {
        double s = 0;
        for (int k = std::max(0, i - 6); k <= j; ++k) s += Bm[(size_t)i * n + k] * Bm[(size_t)j * n + k];
        H[(size_t)i * n + j] = H[(size_t)j * n + i] = s + (i == j ? 1e-3 + (i % 5) : 0.0);
      }
    std::vector<double> L(pbcr_L_doubles(NT) + 16, 0.0), S(pbcr_S_doubles(NT) + 16, 0.0), sc(18 * (NT / 4 + 1), 0.0);
    PbcrMem m{L.data(), S.data(), sc.data(), sc.data() + 6 * (NT / 4 + 1), sc.data() + 12 * (NT / 4 + 1)};
    for (int t = 0; t < Nt; ++t)
      for (int k = 0; k < 6; ++k) {
        const int i = 6 * t + k;
        double *B = pbcr_blk(m.L, t);
        for (int d = 1; d <= 6; ++d) B[6 * k + d - 1] = (i - d >= 0) ? H[(size_t)i * n + i - d] : 0.0;
        B[36 + k] = H[(size_t)i * n + i];
      }
    host_factor(m, Nt, NT);
    // dense Cholesky reference
    std::vector<long double> C((size_t)n * n, 0.0L);
    for (int i = 0; i < n; ++i)
      for (int j = std::max(0, i - 6); j <= i; ++j) {
        long double s = H[(size_t)i * n + j];
        for (int k = std::max(0, i - 6); k < j; ++k) s -= C[(size_t)i * n + k] * C[(size_t)j * n + k];
        C[(size_t)i * n + j] = (i == j) ? sqrtl(s) : s / C[(size_t)j * n + j];
      }
    for (int rep = 0; rep < 3; ++rep) {
      std::vector<double> b(6 * NT, 0.0), tmp(6 * NT, 0.0), b0(n);
      for (int t = 0; t < Nt; ++t)
        for (int k = 0; k < 6; ++k) b0[6 * t + k] = b[k * NT + t] = U(rng) * 10.0;
      host_solve(m, b.data(), tmp.data(), Nt, NT, NT);
      std::vector<long double> y(n), x(n);
      for (int i = 0; i < n; ++i) {
        long double s = b0[i];
        for (int k = std::max(0, i - 6); k < i; ++k) s -= C[(size_t)i * n + k] * y[k];
        y[i] = s / C[(size_t)i * n + i];
      }
      for (int i = n - 1; i >= 0; --i) {
        long double s = y[i];
        for (int k = i + 1; k <= std::min(n - 1, i + 6); ++k) s -= C[(size_t)k * n + i] * x[k];
        x[i] = s / C[(size_t)i * n + i];
      }
      double err = 0, nrm = 0;
      for (int t = 0; t < Nt; ++t)
        for (int k = 0; k < 6; ++k) {
          err = std::fmax(err, std::fabs((double)(b[k * NT + t] - x[6 * t + k])));
          nrm = std::fmax(nrm, std::fabs((double)x[6 * t + k]));
        }
      worst = std::fmax(worst, err / nrm);
      if (!(err / nrm < 1e-9)) { std::printf("FAIL Nt=%d rel err %.3e\n", Nt, err / nrm); return 1; }
    }
  }
  std::printf("pbcr ok: worst relative error %.3e\n", worst);
  return 0;
}
