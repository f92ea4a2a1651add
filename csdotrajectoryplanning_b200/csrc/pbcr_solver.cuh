// Direct solver for the reduced KKT matrix of one agent QP (symmetric positive definite, block
// tridiagonal with 6x6 blocks = one time step, scalar half-bandwidth 6), run by the WHOLE CTA:
// short horizon partitions + block cyclic reduction on the separators ("PBCR").
//
// The Nt time blocks are cut into groups of kPM = 4: kPLen = 3 interior blocks followed by one
// separator block,
//     [i i i][s] [i i i][s] ... [i i (i)]
// (a nested-dissection ordering of the same matrix: still an exact LDL'-type direct solve, only the
// rounding differs from a sequential band factor).
//   factor:  F1  banded LDL' of every 3-block interior            one thread per partition
//            F2  Schur complement of the separators               one thread per partition
//            F3  assemble the separator system S (block tridiagonal, dense 6x6 blocks)
//            F4  block cyclic reduction of S: log2 levels; at stride s every other active separator e is
//                eliminated: Ainv_e = A_e^-1, Wm_e = Ainv_e S[e][e-s], Wp_e = Ainv_e S[e][e+s]; the
//                survivors get A_k -= S[k][e] Ainv_e S[e][k] and the fill S[k][k-2s]
//   solve :  S1  z_p = H_pp^-1 b_p  (forward + backward sweep over 3 blocks)      thread per partition
//            S2  g_j = b_sep(j) - couplings * z                                   thread per (j, row)
//            S3  BCR forward (y_e = Ainv_e g_e; g_k -= Wp_e1' g_e1 + Wm_e2' g_e2) and backward
//                (x_e = y_e - Wm_e x_{e-s} - Wp_e x_{e+s}), one thread per (separator, row): every
//                level is 6..12 dependent FMAs deep
//            S4  x_p = H_pp^-1 (b_p - couplings * x_sep)  (two more 3-block sweeps)  thread per partition
// The dependent chain of a solve is 12 block sweeps + 2 log2(Nt/4) short levels, independent of how the
// horizon compares with the warp size; a 256-step horizon takes the same number of steps as a 96-step one.
//
// Storage.  Block record (kBlkD = 42 doubles): row k of block t keeps l_{i,i-d}, d = 1..6 at [6k + d-1]
// (before the factorization: H_{i,i-d}) and 1/d_i at [36 + k] (before: H_ii).  Slots that reach across a
// partition boundary, and all slots of separator blocks, keep the RAW entries of H, which F2/S2/S4 read.
// A group is 4 records + 2 doubles of padding: the group stride is an odd number of 16-byte slots, so the
// threads of a warp (one partition each) read their records with conflict-free LDS.128.
// Separator record (kSepD = 94 doubles): Ainv packed lower triangle [21] | pad | Wm [36] | Wp [36].
//
// Every phase is a plain function of (thread index, number of threads) so that the same code runs on
// the host for the CPU test (tests/cpp/test_pbcr.cpp emulates the threads of a phase with a loop).
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define CSDO_HD __host__ __device__ __forceinline__
#else
#define CSDO_HD inline
#endif
#ifndef DBG_INIT
#define DBG_INIT() do {} while (0)
#define DBG_ACC(i) do {} while (0)
#endif

namespace csdo {

constexpr int kPLen = 3;                 // interior blocks per partition
constexpr int kPM = kPLen + 1;           // blocks per group (interior + separator)
constexpr int kBlkD = 42;                // doubles per block record
constexpr int kGrpD = kBlkD * kPM + 2;   // doubles per group: 85 16-byte slots (odd)
constexpr int kSepWm = 22, kSepWp = 58;  // offsets of Wm / Wp inside a separator record
constexpr int kSepD = 94;                // doubles per separator record

struct PbcrMem {
  double *L;   // ceil(NT / 4) groups
  double *S;   // NT / 4 separator records
  double *g;   // separator right-hand sides, 6 per separator
  double *y;   // BCR forward results, 6 per separator
  double *xs;  // separator solution, 6 per separator
};

struct PGeom {
  int Nt, NP, Ps, Lv;
  CSDO_HD int part_len(int p) const { const int r = Nt - kPM * p; return r < kPLen ? r : kPLen; }
  CSDO_HD int sep_block(int j) const { return kPM * j + kPLen; }
};
CSDO_HD PGeom pbcr_geom(int Nt) {
  PGeom g;
  g.Nt = Nt;
  g.NP = (Nt + kPM - 1) / kPM;
  g.Ps = Nt / kPM;
  g.Lv = 0;
  while ((1 << g.Lv) <= g.Ps) ++g.Lv;  // floor(log2 Ps) + 1 levels (0 when there is no separator)
  return g;
}
// doubles of L / S storage for a padded horizon NT (multiple of 4)
CSDO_HD int pbcr_L_doubles(int NT) { return ((NT + kPM - 1) / kPM) * kGrpD; }
CSDO_HD int pbcr_S_doubles(int NT) { return (NT / kPM) * kSepD; }

CSDO_HD double *pbcr_blk(double *L, int t) { return L + (t / kPM) * kGrpD + (t % kPM) * kBlkD; }
CSDO_HD const double *pbcr_blk(const double *L, int t) { return L + (t / kPM) * kGrpD + (t % kPM) * kBlkD; }
CSDO_HD int pbcr_q(int a, int b) { return a >= b ? a * (a + 1) / 2 + b : b * (b + 1) / 2 + a; }  // packed symmetric

// ---------------------------------------------------------------------------------------------------
// F1: banded LDL' of the interior blocks of one partition (columns before the partition are ignored)
CSDO_HD void pbcr_interior_factor(double *P, int len) {
  for (int tl = 0; tl < len; ++tl) {
    double *B = P + kBlkD * tl;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      double *Li = B + 6 * k;
      double u[7];
#pragma unroll
      for (int d = 6; d >= 1; --d) {
        const bool inblk = (k - d >= 0);
        double s = 0.0;
        if (inblk || tl > 0) {
          s = Li[d - 1];
          const double *Lj = inblk ? B + 6 * (k - d) : B - kBlkD + 6 * (6 + k - d);
#pragma unroll
          for (int e = 6; e > d; --e)
            if (k - e >= 0 || tl > 0) s -= u[e] * Lj[e - d - 1];
        }
        u[d] = s;
      }
      double dsum = B[36 + k];
#pragma unroll
      for (int d = 6; d >= 1; --d) {
        const bool inblk = (k - d >= 0);
        if (inblk || tl > 0) {
          const double dj = inblk ? B[36 + k - d] : B[-kBlkD + 36 + 6 + k - d];
          const double l = u[d] * dj;
          dsum -= u[d] * l;
          Li[d - 1] = l;
        }
      }
      B[36 + k] = 1.0 / dsum;
    }
  }
}

// one block record into registers: Lr[k][d-1] = l_{i,i-d}, dv[k] = 1/d_i.  SH: the record is in shared
// memory -- 21 explicit 128-bit shared loads (left to itself the compiler split them into 64-bit loads,
// which doubles the shared-memory wavefronts and breaks the conflict-free group stride).
template <bool SH>
CSDO_HD void pbcr_load_block(const double *B, double (&Lr)[6][6], double (&dv)[6]) {
#if defined(__CUDA_ARCH__)
  if (SH) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(B);
#pragma unroll
    for (int k = 0; k < 6; ++k)
#pragma unroll
      for (int h = 0; h < 3; ++h)
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(Lr[k][2 * h]), "=d"(Lr[k][2 * h + 1]) : "r"(a + 16u * (3 * k + h)));
#pragma unroll
    for (int h = 0; h < 3; ++h)
      asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(dv[2 * h]), "=d"(dv[2 * h + 1]) : "r"(a + 16u * (18 + h)));
  } else {
    const double2 *r2 = reinterpret_cast<const double2 *>(B);
#pragma unroll
    for (int k = 0; k < 6; ++k)
#pragma unroll
      for (int h = 0; h < 3; ++h) { const double2 v = r2[3 * k + h]; Lr[k][2 * h] = v.x; Lr[k][2 * h + 1] = v.y; }
#pragma unroll
    for (int h = 0; h < 3; ++h) { const double2 v = r2[18 + h]; dv[2 * h] = v.x; dv[2 * h + 1] = v.y; }
  }
#else
  for (int k = 0; k < 6; ++k)
    for (int d = 0; d < 6; ++d) Lr[k][d] = B[6 * k + d];
  for (int k = 0; k < 6; ++k) dv[k] = B[36 + k];
#endif
}

// forward step of one block: p <- L_tt^-1 (p - L_{t,t-1} prev) (unscaled y), yd <- D^-1 y
CSDO_HD void pbcr_fwd_block(const double (&Lr)[6][6], const double (&dv)[6], double (&p)[6], const double (&prev)[6],
                            double (&yd)[6]) {
#pragma unroll
  for (int k = 0; k < 6; ++k)
#pragma unroll
    for (int d = 6; d > k; --d) p[k] = fma(-Lr[k][d - 1], prev[6 + k - d], p[k]);
#pragma unroll
  for (int j = 0; j < 6; ++j)
#pragma unroll
    for (int k = j + 1; k < 6; ++k) p[k] = fma(-Lr[k][k - j - 1], p[j], p[k]);
#pragma unroll
  for (int k = 0; k < 6; ++k) yd[k] = p[k] * dv[k];
}
// backward step of one block, outer-product order: cur (accumulated right-hand side of this block) becomes
// the block's solution; its contribution is pushed into nxt (the block above)
CSDO_HD void pbcr_bwd_block(const double (&Lr)[6][6], double (&cur)[6], double (&nxt)[6]) {
#pragma unroll
  for (int k = 5; k >= 0; --k) {
    const double xv = cur[k];
#pragma unroll
    for (int d = 1; d <= 6; ++d) {
      if (k - d >= 0) cur[k - d] = fma(-Lr[k][d - 1], xv, cur[k - d]);
      else nxt[6 + k - d] = fma(-Lr[k][d - 1], xv, nxt[6 + k - d]);
    }
  }
}

// S1 / S4 sweeps: out = H_pp^-1 in for blocks t0 .. t0+len-1 (len <= 3) of the partition stored at P.
// in / out are SoA vectors (v[k * NT + t]) and may alias.  The intermediate D^-1 y of the (at most three)
// blocks stays in registers between the forward and the backward sweep.  prev_init (optional): solution
// of the separator block above the partition; the first block's cross-boundary slots hold the raw coupling
// entries, so starting the forward sweep from it subtracts coupling * x_sep from the right-hand side on
// the fly.
template <bool SH>
CSDO_HD void pbcr_interior_solve(const double *P, int len, const double *in, double *out, int t0, int NT,
                                 const double *prev_init) {
  double prev[6] = {0, 0, 0, 0, 0, 0};
  if (prev_init) {
#pragma unroll
    for (int k = 0; k < 6; ++k) prev[k] = prev_init[k];
  }
  double y[kPLen][6];
#pragma unroll
  for (int tl = 0; tl < kPLen; ++tl) {  // forward: L y = b
    if (tl < len) {
      double Lr[6][6], dv[6], p[6];
      pbcr_load_block<SH>(P + kBlkD * tl, Lr, dv);
#pragma unroll
      for (int k = 0; k < 6; ++k) p[k] = in[k * NT + t0 + tl];
      pbcr_fwd_block(Lr, dv, p, prev, y[tl]);
#pragma unroll
      for (int k = 0; k < 6; ++k) prev[k] = p[k];
    }
  }
  double dummy[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int tl = kPLen - 1; tl >= 0; --tl) {  // backward: L' x = D^-1 y
    if (tl < len) {
      double Lr[6][6], dv[6];
      pbcr_load_block<SH>(P + kBlkD * tl, Lr, dv);
      if (tl > 0) pbcr_bwd_block(Lr, y[tl], y[tl - 1]);
      else pbcr_bwd_block(Lr, y[tl], dummy);
#pragma unroll
      for (int k = 0; k < 6; ++k) out[k * NT + t0 + tl] = y[tl][k];
    }
  }
}

// raw coupling H[(t, kr)][(t-1, kc)], kc >= kr, kept in row kr of block record B (block t) at d = 6 + kr - kc
CSDO_HD double pbcr_coupling(const double *B, int kr, int kc) { return B[6 * kr + (6 + kr - kc) - 1]; }

// ---------------------------------------------------------------------------------------------------
// F2: Schur-complement contributions of one partition.  With C = coupling of the first block to the
// previous separator and B = coupling of the next separator to the last block:
//   GCC = C' H_pp^-1 C  -> Sprev[0..20]           (packed; Ainv slot of the previous separator)
//   GBB = B H_pp^-1 B'  -> Sown[kSepWp + 0..20]   (packed)
//   GBC = B H_pp^-1 C   -> Sown[kSepWm + 0..35]   ([next-separator unknown a][previous-separator unknown c])
CSDO_HD void pbcr_schur(const double *P, int len, bool has_prev, bool has_next, double *Sprev, double *Sown) {
  double win[6][6];  // win[a][k]: L^-1 C column a at row k of the last block seen
  if (has_prev) {
    double gcc[21];
#pragma unroll
    for (int q = 0; q < 21; ++q) gcc[q] = 0.0;
    double prev[6][6];
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
      for (int k = 0; k < 6; ++k) prev[a][k] = 0.0;
    for (int tl = 0; tl < len; ++tl) {
      const double *B = P + kBlkD * tl;
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const double *r = B + 6 * k;
        const double di = B[36 + k];
        double ya[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) {
          double s = 0.0;
          if (tl == 0 && a >= k) s = pbcr_coupling(P, k, a);  // C column a: nonzero in the first block only
#pragma unroll
          for (int d = 6; d >= 1; --d) {
            const double yv = (k - d >= 0) ? win[a][k - d] : prev[a][6 + k - d];
            // in the first block the cross-boundary slots hold C itself, not L: they multiply prev == 0
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
#pragma unroll
    for (int q = 0; q < 21; ++q) Sprev[q] = gcc[q];
  }
  if (has_next) {
    const double *Bl = P + kBlkD * (len - 1), *Bs = P + kBlkD * len;  // last interior block, separator
    double z[6][6];  // z[a][k]: L^-1 B' column a restricted to the last block
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        double s = (k >= a) ? pbcr_coupling(Bs, a, k) : 0.0;
        const double *r = Bl + 6 * k;
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
        for (int k = 0; k < 6; ++k) s = fma(z[a][k] * Bl[36 + k], z[b][k], s);
        Sown[kSepWp + q++] = s;
      }
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) {
        double s = 0.0;
        if (has_prev) {
#pragma unroll
          for (int k = 0; k < 6; ++k) s = fma(z[a][k] * Bl[36 + k], win[cc][k], s);
        }
        Sown[kSepWm + a * 6 + cc] = s;
      }
  }
}

// F1 + F2 of partition p
CSDO_HD void pbcr_factor_partition(const PbcrMem &m, const PGeom &g, int p) {
  double *P = m.L + p * kGrpD;
  const int len = g.part_len(p);
  pbcr_interior_factor(P, len);
  const bool has_prev = p > 0, has_next = p < g.Ps;
  pbcr_schur(P, len, has_prev, has_next, has_prev ? m.S + (p - 1) * kSepD : nullptr, m.S + (has_next ? p : 0) * kSepD);
}

// F3: task (separator j, e) with e in [0, 57): e < 21 -> entry q = e of A_j = H_ss - GBB_j - GCC_{j+1}
// (in place over the GCC slot), e >= 21 -> entry e - 21 of Lc_j = S[j][j-1] = -GBC_j (in place).
CSDO_HD void pbcr_assemble_task(const PbcrMem &m, const PGeom &g, int j, int e) {
  double *S = m.S + j * kSepD;
  if (e < 21) {
    int a = 0;
    while ((a + 1) * (a + 2) / 2 <= e) ++a;
    const int b = e - a * (a + 1) / 2;  // a >= b
    const double *Bs = pbcr_blk(m.L, g.sep_block(j));
    double v = (a == b) ? Bs[36 + a] : Bs[6 * a + (a - b) - 1];
    v -= S[kSepWp + e];
    if (j + 1 < g.NP) v -= S[e];
    S[e] = v;
  } else {
    const int idx = kSepWm + (e - 21);
    S[idx] = (j > 0) ? -S[idx] : 0.0;
  }
}

// in-place inverse of a symmetric positive definite 6x6 matrix (Gauss-Jordan, no pivoting), symmetrized
CSDO_HD void pbcr_inv6(double (&M)[6][6]) {
#pragma unroll
  for (int piv = 0; piv < 6; ++piv) {
    const double d = 1.0 / M[piv][piv];
#pragma unroll
    for (int c = 0; c < 6; ++c)
      if (c != piv) M[piv][c] *= d;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      if (r == piv) continue;
      const double f = M[r][piv];
#pragma unroll
      for (int c = 0; c < 6; ++c)
        if (c != piv) M[r][c] = fma(-f, M[piv][c], M[r][c]);
      M[r][piv] = -f * d;
    }
    M[piv][piv] = d;
  }
#pragma unroll
  for (int a = 0; a < 6; ++a)
#pragma unroll
    for (int b = 0; b < a; ++b) {
      const double v = 0.5 * (M[a][b] + M[b][a]);
      M[a][b] = v;
      M[b][a] = v;
    }
}

// F4, level with stride s, separator i (1-based) with i / s odd: eliminate it.
//   Ainv_i, Wm_i = Ainv_i Lc_i (Lc_i = S[i][i-s], found in the Wm slot), Wp_i = Ainv_i Lc_{i+s}'.
// The raw Lc_i is stashed in the Wp slot of the left survivor i - s, whose update needs it.
CSDO_HD void pbcr_bcr_eliminate(const PbcrMem &m, const PGeom &g, int i, int s) {
  double *S = m.S + (i - 1) * kSepD;
  double A[6][6];
#pragma unroll
  for (int a = 0; a < 6; ++a)
#pragma unroll
    for (int b = 0; b <= a; ++b) { A[a][b] = S[a * (a + 1) / 2 + b]; A[b][a] = A[a][b]; }
  pbcr_inv6(A);
#pragma unroll
  for (int a = 0; a < 6; ++a)
#pragma unroll
    for (int b = 0; b <= a; ++b) S[a * (a + 1) / 2 + b] = A[a][b];
  if (i - s >= 1) {
    double *Sl = m.S + (i - s - 1) * kSepD;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      double lc[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) { lc[c] = S[kSepWm + 6 * a + c]; Sl[kSepWp + 6 * a + c] = lc[c]; }
    }
    // Wm = A * Lc, column by column of Lc (kept in registers one column at a time)
#pragma unroll
    for (int b = 0; b < 6; ++b) {
      double col[6], w[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) col[c] = Sl[kSepWp + 6 * c + b];
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        double acc = 0.0;
#pragma unroll
        for (int c = 0; c < 6; ++c) acc = fma(A[a][c], col[c], acc);
        w[a] = acc;
      }
#pragma unroll
      for (int a = 0; a < 6; ++a) S[kSepWm + 6 * a + b] = w[a];
    }
  }
  if (i + s <= g.Ps) {
    const double *Sr = m.S + (i + s - 1) * kSepD;  // right neighbour: its Wm slot holds the raw Lc_{i+s}
#pragma unroll
    for (int b = 0; b < 6; ++b) {  // column b of Lc_{i+s}' = row b of Lc_{i+s}
      double col[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) col[c] = Sr[kSepWm + 6 * b + c];
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        double acc = 0.0;
#pragma unroll
        for (int c = 0; c < 6; ++c) acc = fma(A[a][c], col[c], acc);
        S[kSepWp + 6 * a + b] = acc;
      }
    }
  }
}

// F4, survivor k (1-based, k / s even) of the level with stride s:
//   A_k -= Lc_k Wp_{k-s} + Lc_{k+s}' Wm_{k+s};   Lc_k <- S[k][k-2s] = -Lc_k Wm_{k-s}
CSDO_HD void pbcr_bcr_survive(const PbcrMem &m, const PGeom &g, int k, int s) {
  double *S = m.S + (k - 1) * kSepD;
  const double *S1 = m.S + (k - s - 1) * kSepD;
  double A[21];
#pragma unroll
  for (int q = 0; q < 21; ++q) A[q] = S[q];
  double Lc[6][6];
#pragma unroll
  for (int a = 0; a < 6; ++a)
#pragma unroll
    for (int c = 0; c < 6; ++c) Lc[a][c] = S[kSepWm + 6 * a + c];
#pragma unroll
  for (int b = 0; b < 6; ++b) {  // column b of T1 = Wp_{k-s}
    double col[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) col[c] = S1[kSepWp + 6 * c + b];
#pragma unroll
    for (int a = b; a < 6; ++a) {
      double acc = A[a * (a + 1) / 2 + b];
#pragma unroll
      for (int c = 0; c < 6; ++c) acc = fma(-Lc[a][c], col[c], acc);
      A[a * (a + 1) / 2 + b] = acc;
    }
  }
  if (k + s <= g.Ps) {
    const double *S2 = m.S + (k + s - 1) * kSepD;
    double Le[6][6];  // raw Lc_{k+s}, stashed in this record's Wp slot by the eliminated neighbour
#pragma unroll
    for (int c = 0; c < 6; ++c)
#pragma unroll
      for (int a = 0; a < 6; ++a) Le[c][a] = S[kSepWp + 6 * c + a];
#pragma unroll
    for (int b = 0; b < 6; ++b) {
      double col[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) col[c] = S2[kSepWm + 6 * c + b];
#pragma unroll
      for (int a = b; a < 6; ++a) {
        double acc = A[a * (a + 1) / 2 + b];
#pragma unroll
        for (int c = 0; c < 6; ++c) acc = fma(-Le[c][a], col[c], acc);
        A[a * (a + 1) / 2 + b] = acc;
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 21; ++q) S[q] = A[q];
  const bool has_left = (k - 2 * s >= 1);
#pragma unroll
  for (int b = 0; b < 6; ++b) {
    double col[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) col[c] = has_left ? S1[kSepWm + 6 * c + b] : 0.0;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      double acc = 0.0;
#pragma unroll
      for (int c = 0; c < 6; ++c) acc = fma(-Lc[a][c], col[c], acc);
      S[kSepWm + 6 * a + b] = acc;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// S2, task (separator j, row k): g = b_sep - H[sep][sep-1] z_{sep-1} - H[sep+1][sep]' z_{sep+1}
CSDO_HD void pbcr_sep_rhs_task(const PbcrMem &m, const PGeom &g, const double *b, const double *z, int NT, int j,
                               int k) {
  const int T = g.sep_block(j);
  const double *Bs = pbcr_blk(m.L, T);
  double s = b[k * NT + T];
#pragma unroll
  for (int kc = 0; kc < 6; ++kc)
    if (kc >= k) s = fma(-pbcr_coupling(Bs, k, kc), z[kc * NT + T - 1], s);
  if (T + 1 < g.Nt) {
    const double *Bn = pbcr_blk(m.L, T + 1);
#pragma unroll
    for (int kr = 0; kr < 6; ++kr)
      if (kr <= k) s = fma(-pbcr_coupling(Bn, kr, k), z[kr * NT + T + 1], s);
  }
  m.g[6 * j + k] = s;
}

// S3 forward, level with stride s, task (SURVIVOR separator i = 2 s (ms + 1), row r):
//   g_i -= Wp_{i-s}' g_{i-s} + Wm_{i+s}' g_{i+s}
// (the eliminated separators' y = Ainv g is folded into the backward pass, so every task of a level does
// the same work: no divergence inside a warp)
CSDO_HD void pbcr_bcr_forward_task(const PbcrMem &m, const PGeom &g, int s, int ms, int r) {
  const int i = 2 * s * (ms + 1);
  const double *S1 = m.S + (i - s - 1) * kSepD, *g1 = m.g + 6 * (i - s - 1);
  double acc = m.g[6 * (i - 1) + r], acc2 = 0.0;
#pragma unroll
  for (int c = 0; c < 6; ++c) acc = fma(-S1[kSepWp + 6 * c + r], g1[c], acc);
  if (i + s <= g.Ps) {
    const double *S2 = m.S + (i + s - 1) * kSepD, *g2 = m.g + 6 * (i + s - 1);
#pragma unroll
    for (int c = 0; c < 6; ++c) acc2 = fma(S2[kSepWm + 6 * c + r], g2[c], acc2);
  }
  m.g[6 * (i - 1) + r] = acc - acc2;
}

// S3 backward, level with stride s, task (eliminated separator i = s (2 mo + 1), row r):
//   x_i = Ainv_i g_i - Wm_i x_{i-s} - Wp_i x_{i+s}
CSDO_HD void pbcr_bcr_backward_task(const PbcrMem &m, const PGeom &g, int s, int mo, int r) {
  const int i = s * (2 * mo + 1);
  const double *S = m.S + (i - 1) * kSepD, *gi = m.g + 6 * (i - 1);
  double acc = 0.0, acc1 = 0.0, acc2 = 0.0;
#pragma unroll
  for (int c = 0; c < 6; ++c) acc = fma(S[pbcr_q(r, c)], gi[c], acc);
  if (i - s >= 1) {
    const double *x1 = m.xs + 6 * (i - s - 1);
#pragma unroll
    for (int c = 0; c < 6; ++c) acc1 = fma(S[kSepWm + 6 * r + c], x1[c], acc1);
  }
  if (i + s <= g.Ps) {
    const double *x2 = m.xs + 6 * (i + s - 1);
#pragma unroll
    for (int c = 0; c < 6; ++c) acc2 = fma(S[kSepWp + 6 * r + c], x2[c], acc2);
  }
  m.xs[6 * (i - 1) + r] = acc - (acc1 + acc2);
}

// S4 for partition p: b of the last interior block -= H[sep][last]' x_sep, then the sweeps with the
// previous separator's solution riding on the forward sweep.  b is overwritten by x.
// the sweeps through a function pointer (device only, see pbcr_solve_cta): same arguments as pbcr_interior_solve
using PbcrSweepFn = void (*)(const double *, int, const double *, double *, int, int, const double *);

template <bool SH, bool INDIRECT = false>
CSDO_HD void pbcr_final_partition(const PbcrMem &m, const PGeom &g, double *b, int NT, int p, void *sweep_fn = nullptr) {
  const double *P = m.L + p * kGrpD;
  const int len = g.part_len(p), t0 = kPM * p;
  if (p < g.Ps) {
    const double *Bs = P + kBlkD * len, *xn = m.xs + 6 * p;
    const int tl = t0 + len - 1;
#pragma unroll
    for (int kc = 0; kc < 6; ++kc) {
      double a = b[kc * NT + tl];
#pragma unroll
      for (int kr = 0; kr < 6; ++kr)
        if (kr <= kc) a = fma(-pbcr_coupling(Bs, kr, kc), xn[kr], a);
      b[kc * NT + tl] = a;
    }
  }
  if (INDIRECT) reinterpret_cast<PbcrSweepFn>(sweep_fn)(P, len, b, b, t0, NT, p > 0 ? m.xs + 6 * (p - 1) : nullptr);
  else pbcr_interior_solve<SH>(P, len, b, b, t0, NT, p > 0 ? m.xs + 6 * (p - 1) : nullptr);
}

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------------------------------
// CTA-wide drivers (every thread of the block calls them; they end with a barrier)
// SH: the factor lives in shared memory (it may live in global scratch for very long horizons)
template <bool SH>
__device__ __forceinline__ void pbcr_factor_cta(const PbcrMem &m, int Nt) {
  if (SH) __builtin_assume(__isShared(m.L));
  __builtin_assume(__isShared(m.S));
  const PGeom g = pbcr_geom(Nt);
  const int tid = threadIdx.x, nth = blockDim.x;
  if (tid < g.NP) pbcr_factor_partition(m, g, tid);
  __syncthreads();
  for (int task = tid; task < 57 * g.Ps; task += nth) pbcr_assemble_task(m, g, task / 57, task % 57);
  __syncthreads();
  for (int lv = 0; lv < g.Lv; ++lv) {
    const int s = 1 << lv, nact = g.Ps >> lv;  // active separators: i = s, 2s, ..., nact * s
    const int nel = (nact + 1) >> 1, nsv = nact >> 1;
    if (tid < nel) pbcr_bcr_eliminate(m, g, s * (2 * tid + 1), s);
    __syncthreads();
    if (tid < nsv) pbcr_bcr_survive(m, g, s * (2 * tid + 2), s);
    __syncthreads();
  }
}

// b (SoA, stride NT) is overwritten by H^-1 b; tmp: 6 NT doubles of scratch.  m is taken BY VALUE: the five
// pointers then live in registers (a reference to the shared context made every access a dependent pair of
// loads, re-done after every barrier).  The last BCR levels (at most 10 active separators = 60 row tasks)
// are run by warp 0 alone between warp-level barriers.
// sweep_fn (optional): entry point of an out-of-line copy of pbcr_interior_solve<SH>.  Only the threads that
// own a partition run the sweeps, and only the sweeps want ~200 registers: called through a pointer they get a
// register allocation of their own (full ABI), while the rest of the solve stays in the caller's.
template <bool SH>
__device__ __noinline__ void pbcr_sweep_entry(const double *P, int len, const double *in, double *out, int t0, int NT,
                                              const double *prev_init) {
  if (SH) { __builtin_assume(__isShared(P)); }
  __builtin_assume(__isShared(in)); __builtin_assume(__isShared(out));
  if (prev_init) __builtin_assume(__isShared(prev_init));
  pbcr_interior_solve<SH>(P, len, in, out, t0, NT, prev_init);
}

template <bool SH, bool INDIRECT = false>
__device__ __forceinline__ void pbcr_solve_cta(const PbcrMem m, double *b, double *tmp, int Nt, int NT,
                                               void *sweep_fn = nullptr) {
  if (SH) __builtin_assume(__isShared(m.L));
  __builtin_assume(__isShared(m.S)); __builtin_assume(__isShared(m.g)); __builtin_assume(__isShared(m.y));
  __builtin_assume(__isShared(m.xs)); __builtin_assume(__isShared(b)); __builtin_assume(__isShared(tmp));
  const PGeom g = pbcr_geom(Nt);
  const int tid = threadIdx.x, nth = blockDim.x;
  if (g.Ps == 0) {
    if (tid == 0) {
      if (INDIRECT) reinterpret_cast<PbcrSweepFn>(sweep_fn)(m.L, g.part_len(0), b, b, 0, NT, nullptr);
      else pbcr_interior_solve<SH>(m.L, g.part_len(0), b, b, 0, NT, nullptr);
    }
    __syncthreads();
    return;
  }
  DBG_INIT();
  if (tid < g.NP) {
    if (INDIRECT) reinterpret_cast<PbcrSweepFn>(sweep_fn)(m.L + tid * kGrpD, g.part_len(tid), b, tmp, kPM * tid, NT, nullptr);
    else pbcr_interior_solve<SH>(m.L + tid * kGrpD, g.part_len(tid), b, tmp, kPM * tid, NT, nullptr);
  }
  DBG_ACC(0);   // S1 sweeps (thread 0 takes part)
  __syncthreads();
  DBG_ACC(1);   // barrier after S1
  for (int task = tid; task < 6 * g.Ps; task += nth) pbcr_sep_rhs_task(m, g, b, tmp, NT, task / 6, task % 6);
  __syncthreads();
  DBG_ACC(2);   // S2 + barrier
  int lw = 0;   // levels [0, lw) are wide (whole CTA), [lw, Lv) narrow (warp 0: at most 32 row tasks a level)
  while (lw < g.Lv && 6 * (((g.Ps >> lw) + 1) >> 1) > 32) ++lw;
  for (int lv = 0; lv < lw; ++lv) {
    const int s = 1 << lv, nsv = (g.Ps >> lv) >> 1;
    for (int task = tid; task < 6 * nsv; task += nth) pbcr_bcr_forward_task(m, g, s, task / 6, task % 6);
    __syncthreads();
  }
  DBG_ACC(6);   // BCR wide forward levels
  if (tid < 32) {
    for (int lv = lw; lv < g.Lv; ++lv) {
      const int s = 1 << lv, nsv = (g.Ps >> lv) >> 1;
      if (tid < 6 * nsv) pbcr_bcr_forward_task(m, g, s, tid / 6, tid % 6);
      __syncwarp();
    }
    for (int lv = g.Lv - 1; lv >= lw; --lv) {
      const int s = 1 << lv, nel = ((g.Ps >> lv) + 1) >> 1;
      if (tid < 6 * nel) pbcr_bcr_backward_task(m, g, s, tid / 6, tid % 6);
      __syncwarp();
    }
  }
  DBG_ACC(7);   // BCR narrow levels (warp 0)
  __syncthreads();
  for (int lv = lw - 1; lv >= 0; --lv) {
    const int s = 1 << lv, nel = ((g.Ps >> lv) + 1) >> 1;
    for (int task = tid; task < 6 * nel; task += nth) pbcr_bcr_backward_task(m, g, s, task / 6, task % 6);
    __syncthreads();
  }
  DBG_ACC(3);   // BCR forward + backward levels
  if (tid < g.NP) pbcr_final_partition<SH, INDIRECT>(m, g, b, NT, tid, sweep_fn);
  DBG_ACC(4);   // S4 sweeps
  for (int task = tid; task < 6 * g.Ps; task += nth) b[(task % 6) * NT + g.sep_block(task / 6)] = m.xs[task];
  __syncthreads();
  DBG_ACC(5);   // scatter + final barrier
}
#endif

}  // namespace csdo
