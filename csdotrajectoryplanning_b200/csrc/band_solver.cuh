// Horizon-partitioned LDL' of the reduced KKT matrix (block tridiagonal, 6x6
// blocks, scalar half-bandwidth 6) and the matching solve, run by ONE warp.
//
// The Nt time blocks are cut into P <= 16 partitions separated by P-1 single
// "separator" blocks (a nested-dissection ordering of the same matrix, so the
// solve is still an exact direct solve; only rounding differs from a
// sequential band factor):
//     [interior_0][sep_0][interior_1][sep_1] ... [interior_{P-1}]
// Lane p owns partition p.  The separator system S (block tridiagonal, P-1
// blocks) is dissected once more: for P in {8, 12, 16} every (nb+1)-th
// separator (nb = 1, 2, 3) is a level-2 separator, the 4 groups of nb
// separators between them and the 3 level-2 separators' Schur complement R are
// inverted densely (<= 18 x 18 each); for P <= 4 all separators form R.
//   factor:  F1  banded LDL' of every interior (lanes in lockstep)
//            F2  Schur complement of the separators, streamed per partition
//            F3  level 2: group inverses, R, R^-1 (whole warp, dense)
//   solve :  S1  z_p = H_pp^-1 b_p                (forward + backward sweeps)
//            S2  g = b_sep - coupling * z ; x_sep = S^-1 g  (level-2 mat-vecs)
//            S3  x_p = H_pp^-1 (b_p - coupling * x_sep)      (two more sweeps)
// A sweep step handles one 6x6 block from registers: the part that depends on
// the neighbouring block is 6 independent DFMA chains, the intra-block triangle
// runs in outer-product order, so the loop-carried chain is ~6 DFMAs per block
// and the step is bound by the issue rate of one warp (measured, see
// scripts/micro/sweep_bench.cu) -- hence many short partitions.
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

// (the cycle timers of this solver were removed with the round-2 solver; the macros stay as no-ops)
#define DBG_T(i) do { (void)dbg_t0; } while (0)
#define DBG_TB(i) do { (void)dbg_t0b; } while (0)
#define DBG_CLOCK() 0ll

struct Parts {
  int P, base, rem;
  __device__ __forceinline__ int len(int p) const { return base + (p < rem ? 1 : 0); }
  __device__ __forceinline__ int start(int p) const { return p * (base + 1) + (p < rem ? p : rem); }
  __device__ __forceinline__ int sep(int j) const { return start(j) + len(j); }  // block index of separator j
  // Bank skew (in doubles) of the rows of partition p / separator p: the lanes of the solver warp read
  // their blocks with LDS.128 in lockstep; shifting partition p so that its rows start at 16-byte
  // slot p (mod 8) of the 128-byte bank window makes 8 neighbouring loads conflict-free (a block is
  // 288 B).  The offsets accumulate (each partition is pushed 0..7 slots further than the previous
  // one), so the shifted partitions never overlap.  The table lives in the shared context.
  const int *tab;
  __device__ __forceinline__ int skew(int p) const { return tab[p]; }
  // partition that owns block t (a separator belongs to the partition above it)
  __device__ __forceinline__ int owner(int t) const {
    const int big = rem * (base + 2);
    return t < big ? t / (base + 2) : rem + (t - big) / (base + 1);
  }
  __device__ __forceinline__ int skew_of_block(int t) const { return P == 1 ? 0 : skew(owner(t)); }
};

// P in {1, 2, 3, 4, 8, 12, 16}: at least ~4 interior blocks per partition
__device__ __forceinline__ Parts make_parts(int Nt, const int *tab) {
  Parts q;
  int P = (Nt + 1) / 8;
  P = P < 1 ? 1 : P;
  if (P > 4) P = Nt >= 80 ? 16 : (Nt >= 60 ? 12 : 8);
  q.P = P;
  const int interior = Nt - (P - 1);
  q.base = interior / P;
  q.rem = interior % P;
  q.tab = tab;
  return q;
}
// fills the skew table (kMaxP ints) of a horizon; one thread
__device__ __forceinline__ void fill_skew_table(int Nt, int *tab) {
  const Parts q = make_parts(Nt, tab);
  int acc16 = 0;  // accumulated shift in 16-byte slots
  for (int p = 0; p < kMaxP; ++p) {
    if (p < q.P && q.P > 1) acc16 += (p - (18 * q.start(p) + acc16)) & 7;
    tab[p] = 2 * acc16;
  }
}

// level-2 dissection of the P-1 separators
struct Lvl2 {
  int nb;   // separators per group (0: no groups, all separators are level-2 separators)
  int ns2;  // level-2 separators
  __device__ __forceinline__ int n_g() const { return 6 * nb; }
  __device__ __forceinline__ int n_r() const { return 6 * ns2; }
  __device__ __forceinline__ int sigma(int s) const { return nb ? s * (nb + 1) + nb : s; }  // separator index
};
__device__ __forceinline__ Lvl2 make_lvl2(int P) {
  Lvl2 l;
  const int Ps = P - 1;
  if (Ps <= 3) { l.nb = 0; l.ns2 = Ps; }
  else { l.nb = (Ps - 3) / 4; l.ns2 = 3; }
  return l;
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

// prev_init (optional): solution of the separator block above the partition.  The first block's row
// slots that reach across the partition boundary hold the raw coupling entries, so starting the forward
// sweep from it subtracts coupling * x_sep from the right-hand side on the fly.
__device__ __forceinline__ void interior_solve(const double *__restrict__ L6, const double *__restrict__ dinv,
                                               const double *in, double *out, int t0, int t1, int NT,
                                               const double *prev_init = nullptr) {
  // forward: L y = b, stores y * dinv
  double prev[6] = {0, 0, 0, 0, 0, 0};
  if (prev_init) {
#pragma unroll
    for (int k = 0; k < 6; ++k) prev[k] = prev_init[k];
  }
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
static __device__ void schur_partition(const BandMem &bm, const Parts &pt, int p) {
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

// In-place Gauss-Jordan inversion of an n x n SPD matrix (no pivoting) by a group of `gs` lanes
// (lane index lg in the group); prow = n doubles of scratch.  All lanes of the warp call it together.
__device__ __forceinline__ void gj_invert(double *M, int n, int lg, int gs, double *prow, bool active) {
  for (int piv = 0; piv < n; ++piv) {
    const double d = active ? 1.0 / M[piv * n + piv] : 0.0;
    __syncwarp();
    if (active)
      for (int cidx = lg; cidx < n; cidx += gs) prow[cidx] = (cidx == piv) ? d : M[piv * n + cidx] * d;
    __syncwarp();
    if (active)
      for (int r = lg; r < n; r += gs) {
        if (r == piv) continue;
        double2 *Mr = reinterpret_cast<double2 *>(M + r * n);
        const double2 *pr = reinterpret_cast<const double2 *>(prow);
        const double f = M[r * n + piv];
        for (int c2 = 0; c2 < n / 2; c2 += 3) {  // n is a multiple of 6
          double2 s0 = Mr[c2], s1 = Mr[c2 + 1], s2 = Mr[c2 + 2];
          const double2 p0 = pr[c2], p1 = pr[c2 + 1], p2 = pr[c2 + 2];
          s0.x = fma(-f, p0.x, s0.x); s0.y = fma(-f, p0.y, s0.y);
          s1.x = fma(-f, p1.x, s1.x); s1.y = fma(-f, p1.y, s1.y);
          s2.x = fma(-f, p2.x, s2.x); s2.y = fma(-f, p2.y, s2.y);
          Mr[c2] = s0; Mr[c2 + 1] = s1; Mr[c2 + 2] = s2;
        }
        M[r * n + piv] = -f * d;
      }
    __syncwarp();
    if (active)
      for (int cidx = lg; cidx < n; cidx += gs) M[piv * n + cidx] = prow[cidx];
    __syncwarp();
  }
}

// ---- whole factorization, executed by one warp (all 32 lanes call it) ----
// __noinline__: the factor/solve get their own register allocation instead of competing with the
// register-resident row state of the caller (saved/restored around the call by the solver warp only)
template <bool SH>
__device__ __noinline__ void band_factor_warp(const BandMem bm, int Nt) {
  if (SH) { __builtin_assume(__isShared(bm.L6)); __builtin_assume(__isShared(bm.dinv)); }
  __builtin_assume(__isShared(bm.Sinv)); __builtin_assume(__isShared(bm.sv)); __builtin_assume(__isShared(bm.G));
  __builtin_assume(__isShared(bm.tab));
  const Parts pt = make_parts(Nt, bm.tab);
  const int lane = threadIdx.x & 31;
  long long dbg_t0 = DBG_CLOCK();
  if (lane < pt.P)
    interior_factor(bm.L6 + (pt.P == 1 ? 0 : pt.skew(lane)), bm.dinv, pt.start(lane), pt.start(lane) + pt.len(lane));
  __syncwarp();
  DBG_T(4);
  if (pt.P == 1) return;
  if (lane < pt.P) schur_partition(bm, pt, lane);
  __syncwarp();
  DBG_T(5);
  // F3: level 2.  S has diagonal blocks A_j = H_TT - GBB(partition j) - GCC(partition j+1) and
  // sub-diagonal blocks B_j = S[j][j-1] = -GBC(partition j).
  const Lvl2 l2 = make_lvl2(pt.P);
  const int Ps = pt.P - 1, nb = l2.nb, n_g = l2.n_g(), nR = l2.n_r();
  double *Tinv = bm.Sinv, *R = bm.Sinv + kL2Tinv, *Bc = R + kL2R;
  long long dbg_t0b = DBG_CLOCK();
  for (int e = lane; e < 4 * n_g * n_g; e += 32) Tinv[e] = 0.0;
  for (int e = lane; e < nR * nR; e += 32) R[e] = 0.0;
  __syncwarp();
  for (int e = lane; e < Ps * 36; e += 32) {
    const int j = e / 36, a = (e % 36) / 6, b = e % 6;  // separator j, entry (a, b) of its 6x6 blocks
    const int T = pt.sep(j);
    const int hi = a > b ? a : b, lo = a > b ? b : a;
    double v = (a == b) ? bm.dinv[6 * T + a] : bm.L6[pt.skew(j) + (size_t)(6 * T + hi) * 6 + (hi - lo) - 1];
    const int q = hi * (hi + 1) / 2 + lo;
    v -= bm.G[j * 78 + 21 + q];
    v -= bm.G[(j + 1) * 78 + q];
    const double cpl = j > 0 ? -bm.G[j * 78 + 42 + a * 6 + b] : 0.0;  // B_j[a][b]
    if (nb == 0) {
      R[(6 * j + a) * nR + 6 * j + b] = v;
      if (j > 0) { R[(6 * j + a) * nR + 6 * (j - 1) + b] = cpl; R[(6 * (j - 1) + b) * nR + 6 * j + a] = cpl; }
    } else {
      const int g = j / (nb + 1), i = j % (nb + 1);
      if (i < nb) {  // member i of group g
        double *Tg = Tinv + g * n_g * n_g;
        Tg[(6 * i + a) * n_g + 6 * i + b] = v;
        if (i > 0) { Tg[(6 * i + a) * n_g + 6 * (i - 1) + b] = cpl; Tg[(6 * (i - 1) + b) * n_g + 6 * i + a] = cpl; }
        else if (g > 0) Bc[(2 * (g - 1) + 1) * 36 + a * 6 + b] = cpl;  // first member <-> level-2 separator g-1
      } else {  // level-2 separator g
        R[(6 * g + a) * nR + 6 * g + b] = v;
        Bc[(2 * g) * 36 + a * 6 + b] = cpl;  // level-2 separator g <-> last member of group g
      }
    }
  }
  __syncwarp();
  DBG_TB(7);
  if (nb) {
    // the 4 group inverses, 8 lanes each
    gj_invert(Tinv + (lane >> 3) * n_g * n_g, n_g, lane & 7, 8, bm.sv + 2 * kMaxNs + (lane >> 3) * 18, true);
    DBG_TB(8);
    // R -= couplings * Tinv * couplings
    const int lo6 = n_g - 6;
    for (int e = lane; e < 5 * 36; e += 32) {
      const int blk = e / 36, a = (e % 36) / 6, b = e % 6;
      if (blk < 3) {
        const int s = blk;
        const double *Bl = Bc + (2 * s) * 36, *Br = Bc + (2 * s + 1) * 36;
        const double *Tl = Tinv + s * n_g * n_g, *Tr = Tinv + (s + 1) * n_g * n_g;
        double acc = 0.0;
        for (int cc = 0; cc < 6; ++cc) {
          double t1 = 0.0, t2 = 0.0;
          for (int d = 0; d < 6; ++d) {
            t1 = fma(Tl[(lo6 + cc) * n_g + lo6 + d], Bl[b * 6 + d], t1);
            t2 = fma(Tr[cc * n_g + d], Br[d * 6 + b], t2);
          }
          acc = fma(Bl[a * 6 + cc], t1, acc);
          acc = fma(Br[cc * 6 + a], t2, acc);
        }
        R[(6 * s + a) * nR + 6 * s + b] -= acc;
      } else {
        const int s = blk - 3;  // R[s+1][s] through group s+1
        const double *Bu = Bc + (2 * (s + 1)) * 36, *Bd = Bc + (2 * s + 1) * 36;
        const double *Tm = Tinv + (s + 1) * n_g * n_g;
        double acc = 0.0;
        for (int cc = 0; cc < 6; ++cc) {
          double t1 = 0.0;
          for (int d = 0; d < 6; ++d) t1 = fma(Tm[(lo6 + cc) * n_g + d], Bd[d * 6 + b], t1);
          acc = fma(Bu[a * 6 + cc], t1, acc);
        }
        R[(6 * (s + 1) + a) * nR + 6 * s + b] = -acc;
        R[(6 * s + b) * nR + 6 * (s + 1) + a] = -acc;
      }
    }
    __syncwarp();
    DBG_TB(9);
  }
  gj_invert(R, nR, lane, 32, bm.sv + 2 * kMaxNs, true);
  DBG_TB(10);
  DBG_T(6);
}

// y_G = Tinv_G v_G for the 4 groups: v and y are separator vectors (group G at offset G * 6 (NB + 1))
template <int NB>
__device__ __forceinline__ void group_matvec(const double *Tinv, const double *v, double *y, int lane) {
  constexpr int NG = 6 * NB, ROWS = 4 * NG, TRIPS = (ROWS + 31) / 32, GS = 6 * (NB + 1);
#pragma unroll
  for (int tr = 0; tr < TRIPS; ++tr) {
    const int e = lane + 32 * tr;
    const bool on = e < ROWS;
    const int ee = on ? e : 0;
    const int G = ee / NG, r = ee % NG;
    const double2 *Tr = reinterpret_cast<const double2 *>(Tinv + G * NG * NG + r * NG);
    const double2 *gv = reinterpret_cast<const double2 *>(v + G * GS);
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int c2 = 0; c2 < NG / 2; ++c2) {
      const double2 tv = Tr[c2], gg = gv[c2];
      s0 = fma(tv.x, gg.x, s0);
      s1 = fma(tv.y, gg.y, s1);
    }
    if (on) y[G * GS + r] = s0 + s1;
  }
}

// x_sep = S^-1 g for 4 groups of NB separators and 3 level-2 separators (g, x_sep, z in bm.sv).
// Values move between the steps through shared memory: compact code matters here, the solver warp
// runs this straight-line code once per ADMM iteration and stalls on instruction fetch otherwise.
template <int NB>
__device__ __forceinline__ void lvl2_solve(const BandMem &bm, int lane) {
  constexpr int NG = 6 * NB, GS = 6 * (NB + 1);
  const double *Tinv = bm.Sinv, *Rinv = bm.Sinv + kL2Tinv, *Bc = Rinv + kL2R;
  double *g = bm.sv, *xs = bm.sv + kMaxNs, *z = bm.sv + 2 * kMaxNs;
  // 1. z_G = Tinv_G g_G
  group_matvec<NB>(Tinv, g, z, lane);
  __syncwarp();
  // 2. level-2 right-hand side r = g_sigma - couplings * z (in place)
  if (lane < 18) {
    const int s = lane / 6, a = lane % 6;
    const double *Bl = Bc + (2 * s) * 36, *Br = Bc + (2 * s + 1) * 36;
    const double *zl = z + s * GS + (NG - 6), *zr = z + (s + 1) * GS;
    double r0 = g[s * GS + NG + a], r1 = 0.0;
#pragma unroll
    for (int cc = 0; cc < 6; ++cc) { r0 = fma(-Bl[a * 6 + cc], zl[cc], r0); r1 = fma(-Br[cc * 6 + a], zr[cc], r1); }
    g[s * GS + NG + a] = r0 + r1;
  }
  __syncwarp();
  // 3. level-2 separators: x_sigma = Rinv r
  if (lane < 18) {
    const double2 *Rr = reinterpret_cast<const double2 *>(Rinv + lane * 18);
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const double2 *gv = reinterpret_cast<const double2 *>(g + s * GS + NG);
#pragma unroll
      for (int h = 0; h < 3; ++h) {
        const double2 rv = Rr[3 * s + h], gg = gv[h];
        s0 = fma(rv.x, gg.x, s0);
        s1 = fma(rv.y, gg.y, s1);
      }
    }
    xs[(lane / 6) * GS + NG + lane % 6] = s0 + s1;
  }
  __syncwarp();
  // 4. move the level-2 solution to the groups' right-hand sides (first / last member of a group)
  if (lane < 24) {
    const int G = lane / 6, a = lane % 6;
    double d0 = 0.0, d1 = 0.0;
    if (G > 0) {
      const double *Bd = Bc + (2 * (G - 1) + 1) * 36, *xv = xs + (G - 1) * GS + NG;
#pragma unroll
      for (int b = 0; b < 6; ++b) d0 = fma(Bd[a * 6 + b], xv[b], d0);
    }
    if (G < 3) {
      const double *Bu = Bc + (2 * G) * 36, *xv = xs + G * GS + NG;
#pragma unroll
      for (int b = 0; b < 6; ++b) d1 = fma(Bu[b * 6 + a], xv[b], d1);
    }
    if (NB == 1) {
      g[G * GS + a] -= d0 + d1;
    } else {
      if (G > 0) g[G * GS + a] -= d0;
      if (G < 3) g[G * GS + (NG - 6) + a] -= d1;
    }
  }
  __syncwarp();
  // 5. x_G = Tinv_G g_G
  group_matvec<NB>(Tinv, g, xs, lane);
  __syncwarp();
}

// ---- solve H x = b: b in `rhs` (SoA, overwritten by x), `tmp` is a scratch vector; one warp ----
template <bool SH>
__device__ __forceinline__ void band_solve_body(const BandMem bm, double *rhs, double *tmp, int Nt, int NT) {
  if (SH) { __builtin_assume(__isShared(bm.L6)); __builtin_assume(__isShared(bm.dinv)); }
  __builtin_assume(__isShared(bm.Sinv)); __builtin_assume(__isShared(bm.sv));
  __builtin_assume(__isShared(rhs)); __builtin_assume(__isShared(tmp));
  __builtin_assume(__isShared(bm.tab));
  long long dbg_t0 = DBG_CLOCK();
  const Parts pt = make_parts(Nt, bm.tab);
  const int lane = threadIdx.x & 31;
  if (pt.P == 1) {
    if (lane == 0) interior_solve(bm.L6, bm.dinv, rhs, rhs, 0, Nt, NT);
    __syncwarp();
    return;
  }
  DBG_T(11);
  const int t0 = pt.start(lane < pt.P ? lane : 0), t1 = t0 + pt.len(lane < pt.P ? lane : 0);
  const int sk = pt.skew(lane < pt.P ? lane : 0);
  if (lane < pt.P) interior_solve(bm.L6 + sk, bm.dinv, rhs, tmp, t0, t1, NT);  // S1
  __syncwarp();
  DBG_T(0);
  const int Ns = 6 * (pt.P - 1);
  double *g = bm.sv, *xs = bm.sv + kMaxNs;
  // S2: separator right-hand side g_j = b_T - H[T][T-1] z_{T-1} - H[T][T+1] z_{T+1}, T = sep(j).
  // Lane p first writes b - (coupling to the last block of ITS partition) for the separator below it,
  // then subtracts the coupling to the first block of its partition from the separator above it.
  if (lane < pt.P - 1) {
    const double *Ls = bm.L6 + sk;  // separator t1 shares the skew of partition `lane`
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      double s = rhs[k * NT + t1];
#pragma unroll
      for (int kc = 0; kc < 6; ++kc)
        if (kc >= k) s = fma(-coupling(Ls, t1, k, kc), tmp[kc * NT + t1 - 1], s);
      g[6 * lane + k] = s;
    }
  }
  __syncwarp();
  if (lane > 0 && lane < pt.P) {
    const double *Lp = bm.L6 + sk;
    double zf[6];
#pragma unroll
    for (int kr = 0; kr < 6; ++kr) zf[kr] = tmp[kr * NT + t0];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      double s = g[6 * (lane - 1) + k];
#pragma unroll
      for (int kr = 0; kr < 6; ++kr)
        if (kr <= k) s = fma(-coupling(Lp, t0, kr, k), zf[kr], s);
      g[6 * (lane - 1) + k] = s;
    }
  }
  __syncwarp();
  DBG_T(1);
  // x_sep = S^-1 g through the level-2 dissection
  const Lvl2 l2 = make_lvl2(pt.P);
  if (l2.nb == 0) {
    const int nR = l2.n_r();
    const double *Rinv = bm.Sinv + kL2Tinv;
    if (lane < nR) {
      double s0 = 0.0, s1 = 0.0;
      const double *Rr = Rinv + lane * nR;
      for (int cidx = 0; cidx < nR; cidx += 2) { s0 = fma(Rr[cidx], g[cidx], s0); s1 = fma(Rr[cidx + 1], g[cidx + 1], s1); }
      xs[lane] = s0 + s1;
    }
    __syncwarp();
  } else if (l2.nb == 1) {
    lvl2_solve<1>(bm, lane);
  } else if (l2.nb == 2) {
    lvl2_solve<2>(bm, lane);
  } else {
    lvl2_solve<3>(bm, lane);
  }
  DBG_T(2);
  for (int e = lane; e < Ns; e += 32) rhs[(e % 6) * NT + pt.sep(e / 6)] = xs[e];
  if (lane < pt.P) {  // S3: interiors with the separator solution moved to the right-hand side
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
    // (the coupling of the first block to the previous separator rides on the forward sweep)
    interior_solve(bm.L6 + sk, bm.dinv, rhs, rhs, t0, t1, NT, lane > 0 ? xs + 6 * (lane - 1) : nullptr);
  }
  __syncwarp();
  DBG_T(3);
}
template <bool SH>
__device__ __noinline__ void band_solve_warp(const BandMem bm, double *rhs, double *tmp, int Nt, int NT) {
  band_solve_body<SH>(bm, rhs, tmp, Nt, NT);
}

}  // namespace csdo
