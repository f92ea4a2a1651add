// Horizon-partitioned LDL' of the reduced KKT matrix (block tridiagonal, 6x6
// blocks, scalar half-bandwidth 6) and the matching solve, run by ONE warp.
//
// The Nt time blocks are cut into P <= 8 partitions separated by P-1 single
// "separator" blocks (a nested-dissection ordering of the same matrix, so the
// solve is still an exact direct solve; only rounding differs from a
// sequential band factor):
//     [interior_0][sep_0][interior_1][sep_1] ... [interior_{P-1}]
// Lane p owns partition p.
//   factor:  F1  banded LDL' of every interior (lanes in lockstep)
//            F2  Schur complement of the separators, streamed per partition
//            F3  dense inverse of the (6(P-1))^2 separator system (whole warp)
//   solve :  S1  z_p = H_pp^-1 b_p                (forward + backward sweeps)
//            S2  g = b_sep - coupling * z ; x_sep = Sinv g   (warp mat-vec)
//            S3  x_p = H_pp^-1 (b_p - coupling * x_sep)      (two more sweeps)
// A sweep step handles one 6x6 block from registers: the part that depends on
// the neighbouring block is 6 independent DFMA chains, the intra-block triangle
// runs in outer-product order, so the loop-carried chain is ~6 DFMAs per block
// and the step is bound by the issue rate of one warp (measured, see
// scripts/micro/sweep_bench.cu).
//
// Storage (time-major row i = 6 t + k; every block has 6 rows -- the last time
// step's missing v, w are dummy unknowns with H_ii = 1 and zero coupling):
//   L6[i*6 + d-1]  = l_{i,i-d}, d = 1..6 (before factor: H_{i,i-d}); slots that
//                    reach across a partition boundary keep the RAW coupling
//                    entries H_{i,i-d}, which S2/S3/F2 read
//   dinv[i]        = 1/d_i (before factor: H_ii); separator rows keep H_ii
#pragma once

#include "dsqp_device.cuh"

namespace csdo {

__device__ unsigned long long g_dbg[16];  // developer counters (CSDO_PROFILE)
#define DBG_T(i) do { if (lane == 0) { long long t_ = clock64(); atomicAdd(&g_dbg[i], (unsigned long long)(t_ - dbg_t0)); dbg_t0 = t_; } } while (0)

constexpr int kMaxP = 8;
constexpr int kMaxNs = 6 * (kMaxP - 1);  // separator unknowns

struct Parts {
  int P, base, rem;
  __device__ __forceinline__ int len(int p) const { return base + (p < rem ? 1 : 0); }
  __device__ __forceinline__ int start(int p) const { return p * (base + 1) + (p < rem ? p : rem); }
  __device__ __forceinline__ int sep(int j) const { return start(j) + len(j); }  // block index of separator j
  // Bank skew (in doubles) of the rows of partition p / separator p: the lanes of the solver warp read
  // their blocks with LDS.128 in lockstep; shifting partition p so that its rows start at 16-byte
  // slot p (mod 8) of the 128-byte bank window makes those 8 loads conflict-free (a block is 288 B).
  // The offsets accumulate (each partition is pushed 0..7 slots further than the previous one), so the
  // shifted partitions never overlap.
  int off[8];
  const int *tab = nullptr;  // optional copy of off[] in shared memory (cheap runtime indexing)
  __device__ __forceinline__ int skew(int p) const {
    if (tab) return tab[p];
    int r = 0;  // select chain: keeps off[] in registers
#pragma unroll
    for (int q = 0; q < 8; ++q) r = (q == p) ? off[q] : r;
    return r;
  }
  // partition that owns block t (a separator belongs to the partition above it)
  __device__ __forceinline__ int owner(int t) const {
    const int big = rem * (base + 2);
    return t < big ? t / (base + 2) : rem + (t - big) / (base + 1);
  }
  __device__ __forceinline__ int skew_of_block(int t) const { return P == 1 ? 0 : skew(owner(t)); }
};
constexpr int kSkewPad = 128;  // extra doubles at the end of the L6 area (<= 8 partitions x 14 doubles)

__device__ __forceinline__ Parts make_parts(int Nt) {
  Parts q;
  int P = (Nt + 1) / 8;
  P = P < 1 ? 1 : (P > kMaxP ? kMaxP : P);
  q.P = P;
  const int interior = Nt - (P - 1);
  q.base = interior / P;
  q.rem = interior % P;
  int acc16 = 0;  // accumulated shift in 16-byte slots
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    if (p < P && P > 1) acc16 += (p - (18 * q.start(p) + acc16)) & 7;
    q.off[p] = 2 * acc16;
  }
  return q;
}

// ---- F1: interior banded LDL' of blocks [t0, t1) (columns before 6*t0 are ignored) ----
__device__ __forceinline__ void interior_factor(double *__restrict__ L6, double *__restrict__ dinv, int t0, int t1) {
  const int i0 = 6 * t0;
  for (int i = i0; i < 6 * t1; ++i) {
    double *Li = L6 + (size_t)i * 6;
    double u[7];
#pragma unroll
    for (int d = 6; d >= 1; --d) {
      const int j = i - d;
      double s = 0.0;
      if (j >= i0) {
        s = Li[d - 1];
        const double *Lj = L6 + (size_t)j * 6;
#pragma unroll
        for (int e = 6; e > d; --e)
          if (i - e >= i0) s -= u[e] * Lj[e - d - 1];
      }
      u[d] = s;
    }
    double dsum = dinv[i];
#pragma unroll
    for (int d = 6; d >= 1; --d) {
      const int j = i - d;
      if (j >= i0) {
        const double l = u[d] * dinv[j];
        dsum -= u[d] * l;
        Li[d - 1] = l;
      }
    }
    dinv[i] = 1.0 / dsum;
  }
}

// ---- sweeps: out = H_pp^-1 in for blocks [t0, t1) ----
__device__ __forceinline__ void load_rows(const double *__restrict__ L6, int t, double (&Lr)[6][6]) {
  const double2 *r2 = reinterpret_cast<const double2 *>(L6 + (size_t)36 * t);
#pragma unroll
  for (int k = 0; k < 6; ++k)
#pragma unroll
    for (int h = 0; h < 3; ++h) {
      const double2 v = r2[3 * k + h];
      Lr[k][2 * h] = v.x;
      Lr[k][2 * h + 1] = v.y;
    }
}

__device__ __forceinline__ void interior_solve(const double *__restrict__ L6, const double *__restrict__ dinv,
                                               const double *in, double *out, int t0, int t1, int NT) {
  // forward: L y = b, stores y * dinv
  double prev[6] = {0, 0, 0, 0, 0, 0};
  for (int t = t0; t < t1; ++t) {
    double Lr[6][6], p[6], dv[6];
    load_rows(L6, t, Lr);
#pragma unroll
    for (int k = 0; k < 6; ++k) { p[k] = in[k * NT + t]; dv[k] = dinv[6 * t + k]; }
    // the part that only needs the previous block: row k uses d = k+1..6 -> prev[6+k-d]
#pragma unroll
    for (int k = 0; k < 6; ++k)
#pragma unroll
      for (int d = 6; d > k; --d) p[k] = fma(-Lr[k][d - 1], prev[6 + k - d], p[k]);
    // intra-block triangle, outer-product order
#pragma unroll
    for (int j = 0; j < 6; ++j)
#pragma unroll
      for (int k = j + 1; k < 6; ++k) p[k] = fma(-Lr[k][k - j - 1], p[j], p[k]);
#pragma unroll
    for (int k = 0; k < 6; ++k) { out[k * NT + t] = p[k] * dv[k]; prev[k] = p[k]; }
  }
  // backward: L' x = D^-1 y, outer-product order (every finished x_i updates the 6 rows above it)
  double a_cur[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) a_cur[k] = out[k * NT + t1 - 1];
  for (int t = t1 - 1; t >= t0; --t) {
    double Lr[6][6], a_prev[6];
    load_rows(L6, t, Lr);
    const int ta = t > t0 ? t - 1 : t0;
#pragma unroll
    for (int k = 0; k < 6; ++k) a_prev[k] = out[k * NT + ta];
#pragma unroll
    for (int k = 5; k >= 0; --k) {
      const double xv = a_cur[k];
      out[k * NT + t] = xv;
#pragma unroll
      for (int d = 1; d <= 6; ++d) {
        if (k - d >= 0) a_cur[k - d] = fma(-Lr[k][d - 1], xv, a_cur[k - d]);
        else a_prev[6 + k - d] = fma(-Lr[k][d - 1], xv, a_prev[6 + k - d]);  // discarded when t == t0
      }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) a_cur[k] = a_prev[k];
  }
}

// raw coupling H[(tr,kr)][(tr-1,kc)] stored in row (tr,kr) at d = 6 + kr - kc (kc >= kr)
__device__ __forceinline__ double coupling(const double *L6, int tr, int kr, int kc) {
  return L6[(size_t)(6 * tr + kr) * 6 + (6 + kr - kc) - 1];
}

// ---- F2: Schur complement contributions of partition p, streamed ----
__device__ void schur_partition(const BandMem &bm, const Parts &pt, int p) {
  const double *L6 = bm.L6 + pt.skew(p), *dinv = bm.dinv;  // own rows and the next separator (same skew)
  const int t0 = pt.start(p), t1 = t0 + pt.len(p);
  double *GCC = bm.G + p * 78, *GBB = GCC + 21, *GBC = GBB + 21;
  double win[6][6];  // win[a][k]: y of column a (coupling to previous separator unknown a) at the last block seen
  double gcc[21];
#pragma unroll
  for (int q = 0; q < 21; ++q) gcc[q] = 0.0;
  const bool has_prev = p > 0;
  if (has_prev) {
    double prev[6][6];
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
      for (int k = 0; k < 6; ++k) prev[a][k] = 0.0;
    for (int t = t0; t < t1; ++t) {
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const double *r = L6 + (size_t)(6 * t + k) * 6;
        const double di = dinv[6 * t + k];
        double ya[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) {
          double s = 0.0;
          if (t == t0 && a >= k) s = coupling(L6, t0, k, a);  // C column a, nonzero in the first block only
#pragma unroll
          for (int d = 6; d >= 1; --d) {
            const double yv = (k - d >= 0) ? win[a][k - d] : prev[a][6 + k - d];
            s = fma(-r[d - 1], yv, s);
          }
          win[a][k] = s;
          ya[a] = s;
        }
        int q = 0;
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
          for (int b = 0; b <= a; ++b) { gcc[q] = fma(ya[a] * di, ya[b], gcc[q]); ++q; }
      }
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int k = 0; k < 6; ++k) prev[a][k] = win[a][k];
    }
  }
#pragma unroll
  for (int q = 0; q < 21; ++q) GCC[q] = gcc[q];
  if (p < pt.P - 1) {
    // B columns: coupling of the next separator (block T) to the last interior block T-1
    const int T = t1, tl = t1 - 1;
    double z[6][6];  // z[a][k]: L^-1 B' column a restricted to the last block
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        double s = (k >= a) ? coupling(L6, T, a, k) : 0.0;
        const double *r = L6 + (size_t)(6 * tl + k) * 6;
#pragma unroll
        for (int d = 1; d <= 5; ++d)
          if (k - d >= 0) s = fma(-r[d - 1], z[a][k - d], s);
        z[a][k] = s;
      }
    int q = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
      for (int b = 0; b <= a; ++b) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) s = fma(z[a][k] * dinv[6 * tl + k], z[b][k], s);
        GBB[q++] = s;
      }
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) {
        double s = 0.0;
        if (has_prev) {
#pragma unroll
          for (int k = 0; k < 6; ++k) s = fma(z[a][k] * dinv[6 * tl + k], win[cc][k], s);
        }
        GBC[a * 6 + cc] = s;  // (next separator unknown a) x (previous separator unknown cc)
      }
  }
}

// ---- whole factorization, executed by one warp (all 32 lanes call it) ----
// __noinline__: the factor/solve get their own register allocation instead of competing with the
// register-resident row state of the caller (saved/restored around the call by the solver warp only)
template <bool SH>
__device__ __noinline__ void band_factor_warp(const BandMem bm, int Nt) {
  if (SH) { __builtin_assume(__isShared(bm.L6)); __builtin_assume(__isShared(bm.dinv)); }
  __builtin_assume(__isShared(bm.Sinv)); __builtin_assume(__isShared(bm.sv)); __builtin_assume(__isShared(bm.G));
  Parts pt = make_parts(Nt);
  const int lane = threadIdx.x & 31;
  {
    int *tab = reinterpret_cast<int *>(bm.sv + 3 * kMaxNs);
    if (lane < 8) tab[lane] = pt.skew(lane);
    __syncwarp();
    pt.tab = tab;
  }
  long long dbg_t0 = clock64();
  if (lane < pt.P)
    interior_factor(bm.L6 + (pt.P == 1 ? 0 : pt.skew(lane)), bm.dinv, pt.start(lane), pt.start(lane) + pt.len(lane));
  __syncwarp();
  DBG_T(4);
  if (pt.P == 1) return;
  if (lane < pt.P) schur_partition(bm, pt, lane);
  __syncwarp();
  DBG_T(5);
  // F3: assemble S (dense, symmetric) and invert it in place (Gauss-Jordan, SPD: no pivoting)
  const int Ns = 6 * (pt.P - 1);
  double *S = bm.Sinv;
  for (int e = lane; e < Ns * Ns; e += 32) S[e] = 0.0;
  __syncwarp();
  for (int e = lane; e < (pt.P - 1) * 36; e += 32) {
    const int j = e / 36, a = (e % 36) / 6, b = e % 6;  // separator j, entry (a, b) of its 6x6 blocks
    const int T = pt.sep(j);
    // diagonal block: H_TT - GBB(partition j) - GCC(partition j+1)
    const int hi = a > b ? a : b, lo = a > b ? b : a;
    double v = (a == b) ? bm.dinv[6 * T + a] : bm.L6[pt.skew(j) + (size_t)(6 * T + hi) * 6 + (hi - lo) - 1];
    const int q = hi * (hi + 1) / 2 + lo;
    v -= bm.G[j * 78 + 21 + q];
    v -= bm.G[(j + 1) * 78 + q];
    S[(6 * j + a) * Ns + 6 * j + b] = v;
    if (j > 0) {  // coupling to the previous separator through partition j: -GBC(partition j)
      const double c = -bm.G[j * 78 + 42 + a * 6 + b];
      S[(6 * j + a) * Ns + 6 * (j - 1) + b] = c;
      S[(6 * (j - 1) + b) * Ns + 6 * j + a] = c;
    }
  }
  __syncwarp();
  // In-place Gauss-Jordan inversion (SPD: no pivoting).  Lane r updates rows r and r+32 with 16-byte
  // shared-memory accesses, 6 columns per trip so the loads of a trip are in flight together; the pivot
  // column is patched after the sweep instead of being tested inside it.  Ns is a multiple of 6.
  double *prow = bm.sv + 2 * kMaxNs;
  for (int piv = 0; piv < Ns; ++piv) {
    const double d = 1.0 / S[piv * Ns + piv];
    __syncwarp();
    for (int cidx = lane; cidx < Ns; cidx += 32) prow[cidx] = (cidx == piv) ? d : S[piv * Ns + cidx] * d;
    __syncwarp();
    for (int r = lane; r < Ns; r += 32) {
      if (r == piv) continue;
      double2 *Sr = reinterpret_cast<double2 *>(S + r * Ns);
      const double2 *pr = reinterpret_cast<const double2 *>(prow);
      const double f = S[r * Ns + piv];
      for (int c2 = 0; c2 < Ns / 2; c2 += 3) {
        double2 s0 = Sr[c2], s1 = Sr[c2 + 1], s2 = Sr[c2 + 2];
        const double2 p0 = pr[c2], p1 = pr[c2 + 1], p2 = pr[c2 + 2];
        s0.x = fma(-f, p0.x, s0.x); s0.y = fma(-f, p0.y, s0.y);
        s1.x = fma(-f, p1.x, s1.x); s1.y = fma(-f, p1.y, s1.y);
        s2.x = fma(-f, p2.x, s2.x); s2.y = fma(-f, p2.y, s2.y);
        Sr[c2] = s0; Sr[c2 + 1] = s1; Sr[c2 + 2] = s2;
      }
      S[r * Ns + piv] = -f * d;
    }
    __syncwarp();
    for (int cidx = lane; cidx < Ns; cidx += 32) S[piv * Ns + cidx] = prow[cidx];
    __syncwarp();
  }
  DBG_T(6);
}

// ---- solve H x = b: b in `rhs` (SoA, overwritten by x), `tmp` is a scratch vector; one warp ----
template <bool SH>
__device__ __noinline__ void band_solve_warp(const BandMem bm, double *rhs, double *tmp, int Nt, int NT) {
  if (SH) { __builtin_assume(__isShared(bm.L6)); __builtin_assume(__isShared(bm.dinv)); }
  __builtin_assume(__isShared(bm.Sinv)); __builtin_assume(__isShared(bm.sv));
  __builtin_assume(__isShared(rhs)); __builtin_assume(__isShared(tmp));
  Parts pt = make_parts(Nt);
  const int lane = threadIdx.x & 31;
  {
    int *tab = reinterpret_cast<int *>(bm.sv + 3 * kMaxNs);
    if (lane < 8) tab[lane] = pt.skew(lane);
    __syncwarp();
    pt.tab = tab;
  }
  if (pt.P == 1) {
    if (lane == 0) interior_solve(bm.L6, bm.dinv, rhs, rhs, 0, Nt, NT);
    __syncwarp();
    return;
  }
  long long dbg_t0 = clock64();
  const int t0 = pt.start(lane < pt.P ? lane : 0), t1 = t0 + pt.len(lane < pt.P ? lane : 0);
  const int sk = pt.skew(lane < pt.P ? lane : 0);
  if (lane < pt.P) interior_solve(bm.L6 + sk, bm.dinv, rhs, tmp, t0, t1, NT);  // S1
  __syncwarp();
  DBG_T(0);
  const int Ns = 6 * (pt.P - 1);
  double *g = bm.sv, *xs = bm.sv + kMaxNs;
  for (int e = lane; e < Ns; e += 32) {  // S2: separator right-hand side
    const int j = e / 6, k = e % 6, T = pt.sep(j);
    double s = rhs[k * NT + T];
#pragma unroll
    for (int kc = 0; kc < 6; ++kc)
      if (kc >= k) s = fma(-coupling(bm.L6 + pt.skew(j), T, k, kc), tmp[kc * NT + T - 1], s);
#pragma unroll
    for (int kr = 0; kr < 6; ++kr)
      if (kr <= k) s = fma(-coupling(bm.L6 + pt.skew(j + 1), T + 1, kr, k), tmp[kr * NT + T + 1], s);
    g[e] = s;
  }
  __syncwarp();
  DBG_T(1);
  for (int r = lane; r < Ns; r += 32) {  // x_sep = Sinv g
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    const double *Sr = bm.Sinv + r * Ns;
#pragma unroll 2
    for (int cidx = 0; cidx < Ns; cidx += 6) {  // Ns is a multiple of 6
      s0 = fma(Sr[cidx], g[cidx], s0); s1 = fma(Sr[cidx + 1], g[cidx + 1], s1); s2 = fma(Sr[cidx + 2], g[cidx + 2], s2);
      s0 = fma(Sr[cidx + 3], g[cidx + 3], s0); s1 = fma(Sr[cidx + 4], g[cidx + 4], s1); s2 = fma(Sr[cidx + 5], g[cidx + 5], s2);
    }
    xs[r] = (s0 + s1) + s2;
  }
  __syncwarp();
  DBG_T(2);
  for (int e = lane; e < Ns; e += 32) rhs[(e % 6) * NT + pt.sep(e / 6)] = xs[e];
  if (lane < pt.P) {  // S3: interiors with the separator solution moved to the right-hand side
    if (lane > 0) {  // rows of the first block couple to the previous separator
      const double *xp = xs + 6 * (lane - 1);
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        double a = rhs[k * NT + t0];
#pragma unroll
        for (int kc = 0; kc < 6; ++kc)
          if (kc >= k) a = fma(-coupling(bm.L6 + sk, t0, k, kc), xp[kc], a);
        rhs[k * NT + t0] = a;
      }
    }
    if (lane < pt.P - 1) {  // columns of the last block couple to the next separator's rows
      const double *xn = xs + 6 * lane;
#pragma unroll
      for (int kc = 0; kc < 6; ++kc) {
        double a = rhs[kc * NT + t1 - 1];
#pragma unroll
        for (int kr = 0; kr < 6; ++kr)
          if (kr <= kc) a = fma(-coupling(bm.L6 + sk, t1, kr, kc), xn[kr], a);
        rhs[kc * NT + t1 - 1] = a;
      }
    }
    interior_solve(bm.L6 + sk, bm.dinv, rhs, rhs, t0, t1, NT);
  }
  __syncwarp();
  DBG_T(3);
}

}  // namespace csdo
