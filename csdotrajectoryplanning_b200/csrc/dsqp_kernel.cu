// DSQP refine kernel for B200 (sm_100a): persistent CTAs, one agent per CTA at
// a time, the agent's whole SQP loop on the device.  See dsqp_device.cuh for the
// data layout.  Follows the reference's iterate sequence in FP64:
//   SolverDSQP::calcIndividualSQP   sqp/dsqp_solver.cc:36-269
//   solveOSQP -> OSQP 0.6.x ADMM    sqp/dsqp_solver.cc:423-555 (third-party solver,
//                                   restated; see oracle/osqp_restate.c for the
//                                   function-by-function correspondence)
//   isFeasible / updateCorridor     sqp/dsqp_solver.cc:292-420, 818-872
//   generateBox & co                sqp/corridor.cc:25-324
//
// This file is compiled TWICE (see the Makefile): -DCSDO_TU=1 holds the kernel variants for block sizes <= 96
// (one-warp band solver, band_solver.cuh), -DCSDO_TU=2 the variants above (CTA-wide solver, pbcr_solver.cuh)
// plus the small kernels and the host-side launchers.  Two modules because ptxas settles the calling
// convention of the out-of-line phases per module: with both solvers' entry points address-taken in one
// module, the 255-register kernels' hot row pass went from 64 to 184 B of register saves (-4 % at Nt = 256).
#if !defined(CSDO_TU) || CSDO_TU < 1 || CSDO_TU > 2
#error "compile with -DCSDO_TU=1 and -DCSDO_TU=2 (see the Makefile)"
#endif
#include "dsqp_device.cuh"
#include "band_solver.cuh"
#include "dsqp_launch.h"
#include <cstdlib>

namespace csdo {

// ===================================================================
// corridor boxes (sqp/corridor.cc) -- adds/compares only, bit-exact
// ===================================================================
struct Box { double x_min, y_min, x_max, y_max; };

struct ObsView {
  const double *obs;  // [No][3] (staged in shared memory by the refine kernel when it fits)
  int No;
  double rv;
  short *cand;        // kMaxCand candidate indices for the calling thread
};

// isBoxValid :252-272 restricted to the candidate list (cand == nullptr: all)
__device__ __forceinline__ bool box_valid(const Box &b, const ObsView &ov, const short *cand, int ncand,
                                          double dimx, double dimy) {
  const double rv = ov.rv;
  if (b.x_min < rv || b.x_max > dimx - rv || b.y_min < rv || b.y_max > dimy - rv) return false;
  const int n = cand ? ncand : ov.No;
  for (int q = 0; q < n; ++q) {
    const int o = cand ? cand[q] : q;
    const double ox = ov.obs[3 * o], oy = ov.obs[3 * o + 1], R = ov.obs[3 * o + 2] + rv;
    // Box::ExpandBox x4 (corridor.h:26-49): y_max+R, x_min-R, y_min-R, x_max+R
    if (b.x_min - R < ox && ox < b.x_max + R && b.y_min - R < oy && oy < b.y_max + R) return false;
  }
  return true;
}

constexpr int kMaxCand = 40;

// generateLocalBox :278-324.  Only obstacles that can ever touch a box grown
// from (xc,yc) are kept as candidates (a conservative superset, so the
// sequence of accepted expansions is unchanged).
static __device__ bool local_box(double xc, double yc, const ObsView &ov, double dimx, double dimy,
                          const csdo_params &P, Box &res) {
  short *cand = ov.cand;  // per-thread slice (shared memory inside the refine kernel)
  int ncand = 0;
  bool overflow = false;
  const double reach = P.box_limit + 2.0 * P.box_ds + 1e-6;
  for (int o = 0; o < ov.No; ++o) {
    const double R = ov.obs[3 * o + 2] + ov.rv + reach;
    if (fabs(ov.obs[3 * o] - xc) < R && fabs(ov.obs[3 * o + 1] - yc) < R) {
      if (ncand < kMaxCand) cand[ncand++] = (short)o;
      else overflow = true;
    }
  }
  const short *cl = overflow ? nullptr : cand;
  int id[4] = {0, 1, 2, 3};
  double lens[4] = {0, 0, 0, 0};
  Box box = {xc, yc, xc, yc};
  int num_expand = 0, n_valid = 4;
  while (n_valid > 0) {
    // Rounds that cannot fail are taken without checks: `gap` is the smallest growth of any live side
    // that could reach a map bound or pull an obstacle centre into the inflated box, so fewer than
    // gap/ds rounds leave every trial valid.  The box sides are still advanced by the same repeated
    // += / -= ds, so the result is bit-identical to the reference's checked expansion.
    {
      double gap = 1e300;
      if (id[0] != -1) gap = fmin(gap, (dimy - ov.rv) - box.y_max);
      if (id[1] != -1) gap = fmin(gap, box.x_min - ov.rv);
      if (id[2] != -1) gap = fmin(gap, box.y_min - ov.rv);
      if (id[3] != -1) gap = fmin(gap, (dimx - ov.rv) - box.x_max);
      const int nc = cl ? ncand : ov.No;
      for (int q = 0; q < nc; ++q) {
        const int o = cl ? cl[q] : q;
        const double ox = ov.obs[3 * o], oy = ov.obs[3 * o + 1], R = ov.obs[3 * o + 2] + ov.rv;
        // growth needed before the centre lies strictly inside the inflated box, per axis; a side that
        // has retired cannot bring the box any closer to an obstacle lying beyond it
        double nx = 0.0, ny = 0.0;
        if (ox >= box.x_max + R) nx = (id[3] != -1) ? ox - (box.x_max + R) : 1e300;
        else if (ox <= box.x_min - R) nx = (id[1] != -1) ? (box.x_min - R) - ox : 1e300;
        if (oy >= box.y_max + R) ny = (id[0] != -1) ? oy - (box.y_max + R) : 1e300;
        else if (oy <= box.y_min - R) ny = (id[2] != -1) ? (box.y_min - R) - oy : 1e300;
        gap = fmin(gap, fmax(nx, ny));
      }
      int safe = (int)floor((gap - 1e-9) / P.box_ds) - 1;
      for (; safe > 0 && n_valid > 0; --safe) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (id[k] == -1) continue;
          if (k == 0) box.y_max += P.box_ds;
          else if (k == 1) box.x_min -= P.box_ds;
          else if (k == 2) box.y_min -= P.box_ds;
          else box.x_max += P.box_ds;
          num_expand++;
          lens[k] += P.box_ds;
          if (lens[k] >= P.box_limit) { n_valid--; id[k] = -1; }
        }
      }
      if (n_valid <= 0) break;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (id[k] == -1) continue;
      Box tr = box;
      if (k == 0) tr.y_max += P.box_ds;
      else if (k == 1) tr.x_min -= P.box_ds;
      else if (k == 2) tr.y_min -= P.box_ds;
      else tr.x_max += P.box_ds;
      if (box_valid(tr, ov, cl, ncand, dimx, dimy)) {
        num_expand++;
        lens[k] += P.box_ds;
        box = tr;
        if (lens[k] >= P.box_limit) { n_valid--; id[k] = -1; }
      } else {
        n_valid--; id[k] = -1;
      }
    }
  }
  res = box;
  return num_expand > 0;
}

// generateBox :124-159 (+ isPointOutOfMap :25-30, projectNearBorder :54-81,
// isPointCollision :32-52, generateLegalPoint :84-122)
static __device__ void generate_box(double x, double y, const ObsView &ov, double dimx, double dimy,
                             const csdo_params &P, Box &box, int &success, int &initial) {
  const double rv = ov.rv;
  initial = 0;
  box = Box{x, y, x, y};
  if (x < rv || x > dimx - rv || y < rv || y > dimy - rv) {
    initial = 1;
    const double eps = 1e-3;
    if (x < rv) x = rv + eps;
    else if (x > dimx - rv) x = dimx - rv - eps;
    if (y < rv) y = rv + eps;
    else if (y > dimy - rv) y = dimy - rv - eps;
  }
  int oc = -1;
  for (int o = 0; o < ov.No; ++o) {  // first hit in container order
    const double ox = ov.obs[3 * o], oy = ov.obs[3 * o + 1], R = ov.obs[3 * o + 2] + rv;
    if (x - R < ox && ox < x + R && y - R < oy && oy < y + R) { oc = o; break; }
  }
  if (oc < 0) {
    success = local_box(x, y, ov, dimx, dimy, P, box) ? 1 : 0;
    return;
  }
  initial = 2;
  const double ocx = ov.obs[3 * oc], ocy = ov.obs[3 * oc + 1], ocr = ov.obs[3 * oc + 2];
  const double theta0 = atan2(y - ocy, x - ocx);
  const double d = rv + ocr + 0.2;
  const int n_cand = 20;
  for (int i = 0; i < n_cand; ++i) {
    int j = i / 2;
    if (i % 2 == 1) j = -j;
    const double theta = __dadd_rn(theta0, j * 2 * M_PI / n_cand);
    x = __dadd_rn(ocx, __dmul_rn(d, cos(theta)));
    y = __dadd_rn(ocy, __dmul_rn(d, sin(theta)));
    if (x > rv && x < dimx - rv && y > rv && y < dimy - rv) {
      Box b = {0, 0, 0, 0};
      local_box(x, y, ov, dimx, dimy, P, b);
      if (box_valid(b, ov, nullptr, 0, dimx, dimy)) { box = b; success = 1; return; }
    }
  }
  box = Box{x, y, x, y};
  success = 0;
}

// calcCorridors (float centres) / updateCorridor (double centres) for the
// agent of this CTA: 2 boxes per step, strided over the CTA's threads.
static __device__ int agent_corridors(const Ctx &c, const csdo_params &P, const double *xs, const double *ys,
                               const double *yaws, double *stage, int stage_doubles, bool double_centres,
                               int *box_status) {
  // Stage the instance's obstacles and the per-thread candidate lists in shared memory that is idle
  // while corridors are generated (`stage`): global/local memory would cost an L2 round trip per access.
  short cand_local[kMaxCand];
  ObsView ov{c.obs(), c.No(), P.rv, cand_local};
  const int cand_doubles = (c.nthr() * kMaxCand * (int)sizeof(short) + 7) / 8;
  if (stage && cand_doubles <= stage_doubles) {
    ov.cand = reinterpret_cast<short *>(stage) + c.tid() * kMaxCand;
    double *so = stage + cand_doubles;
    if (3 * c.No() <= stage_doubles - cand_doubles) {
      for (int i = c.tid(); i < 3 * c.No(); i += c.nthr()) so[i] = c.obs()[i];
      ov.obs = so;
    }
    __syncthreads();
  }
  int illegal = 0;
  const int Nt = c.Nt();
  for (int b = c.tid(); b < 2 * Nt; b += c.nthr()) {
    const int t = b >> 1, rear = b & 1;
    const double off = rear ? P.r2x : P.f2x;
    // separate multiply and add (the reference build has no FMA contraction)
    double cx = __dadd_rn(xs[t], __dmul_rn(off, cos(yaws[t])));
    double cy = __dadd_rn(ys[t], __dmul_rn(off, sin(yaws[t])));
    if (!double_centres) {  // State members are float: motion_planning.h:115-118,230
      cx = (double)(float)cx;
      cy = (double)(float)cy;
    }
    Box bx; int ok, init;
    generate_box(cx, cy, ov, c.dimx(), c.dimy(), P, bx, ok, init);
    if (init > 0) illegal = 1;
    const int base = rear ? 4 : 0;
    c.corr()[(base + 0) * Nt + t] = bx.x_min;
    c.corr()[(base + 1) * Nt + t] = bx.x_max;
    c.corr()[(base + 2) * Nt + t] = bx.y_min;
    c.corr()[(base + 3) * Nt + t] = bx.y_max;
    if (box_status) { box_status[2 * b] = ok; box_status[2 * b + 1] = init; }
  }
  return illegal;
}

// ===================================================================
// row functors
// ===================================================================
template <int NC>
__device__ __forceinline__ double row_dot(const double (&v)[10], int i0, double c0, int i1, double c1, int i2,
                                          double c2, int i3, double c3) {
  double s = c0 * v[i0];
  if (NC > 1) s += c1 * v[i1];
  if (NC > 2) s += c2 * v[i2];
  if (NC > 3) s += c3 * v[i3];
  return s;
}
template <int NC>
__device__ __forceinline__ void row_scatter(double (&acc)[10], double g, int i0, double c0, int i1, double c1,
                                            int i2, double c2, int i3, double c3) {
  acc[i0] += c0 * g;
  if (NC > 1) acc[i1] += c1 * g;
  if (NC > 2) acc[i2] += c2 * g;
  if (NC > 3) acc[i3] += c3 * g;
}

// ADMM row update (OSQP update_xz_tilde/update_z/update_y folded on the single
// state w = z_hat + y/rho: z = clip(w), y = rho (w - z)).
//   MODE 0: warm start (osqp_warm_start_x: z = A x, y = 0)
//   MODE 1: first iteration (z_prev is the unprojected A x)
//   MODE 2: regular iteration
//   MODE 3: re-base after a rho update (keeps z and y, osqp_update_rho)
template <int MODE>
struct StepF {
  static constexpr bool kWriteW = true, kWriteE = false;
  static constexpr int kPlaneOut = 3;  // visit_planes: A' (rho z - y) on x, y, yaw of the plane's step
  __device__ __forceinline__ void plane_begin(const double (&in)[3]) {
    xv[VX] = in[0]; xv[VY] = in[1]; xv[VP] = in[2];
    acc[VX] = 0.0; acc[VY] = 0.0; acc[VP] = 0.0;
  }
  __device__ __forceinline__ void plane_emit(double *o) const { o[0] = acc[VX]; o[1] = acc[VY]; o[2] = acc[VP]; }
  __device__ __forceinline__ void plane_absorb(const double *o) { acc[VX] += o[0]; acc[VY] += o[1]; acc[VP] += o[2]; }
  __device__ __forceinline__ void plane_merge(const StepF &) {}
  double xv[10];
  double acc[10];
  double alpha, rho, rho_old;
  bool store_dy;
  double *dy_base;  // delta_y of this thread's fixed rows (only kept when the agent has no planes)
  int dy_stride;
  template <int NC, bool EQ = false>
  __device__ __forceinline__ void row(int rid, int i0, double c0, int i1, double c1, int i2, double c2, int i3,
                                      double c3, double l, double u, double &w, double &E) {
    const double e = E;
    const double ls = e * l, us = EQ ? ls : e * u;
    // equality rows: us - ls == 0 < kRhoTol and clip(v, ls, ls) == ls for every v
    const double rho_i = EQ ? kRhoEqOverIneq * rho : row_rho(ls, us, rho);
    double wn, zn;
    if (MODE == 0) {
      wn = e * row_dot<NC>(xv, i0, c0, i1, c1, i2, c2, i3, c3);
      zn = wn;
    } else if (MODE == 3) {
      const double wo = w, z = EQ ? ls : clipd(wo, ls, us);
      const double rho_i_old = EQ ? kRhoEqOverIneq * rho_old : row_rho(ls, us, rho_old);
      wn = z + (rho_i_old / rho_i) * (wo - z);
      zn = z;
    } else {
      const double zt = e * row_dot<NC>(xv, i0, c0, i1, c1, i2, c2, i3, c3);
      const double wo = w;
      const double zo = (MODE == 1) ? wo : (EQ ? ls : clipd(wo, ls, us));
      const double yor = wo - zo;  // y_prev / rho
      const double zhat = alpha * zt + (1.0 - alpha) * zo;
      wn = zhat + yor;
      zn = EQ ? ls : clipd(wn, ls, us);
      if (store_dy && rid >= 0) dy_base[rid * dy_stride] = rho_i * (zhat - zn);
    }
    w = wn;
    const double g = e * (rho_i * (2.0 * zn - wn));  // E (rho z - y)
    row_scatter<NC>(acc, g, i0, c0, i1, c1, i2, c2, i3, c3);
  }
};

// residuals / norms (OSQP update_info, compute_pri_tol, compute_dua_tol,
// is_primal_infeasible, compute_rho_estimate)
enum Norm : int {
  N_PRI_U = 0, N_Z_U, N_AX_U, N_PRI_S, N_Z_S, N_AX_S,  // rows
  N_DUA_U, N_PX_U, N_ATY_U, N_DUA_S, N_PX_S, N_ATY_S,   // unknowns
  N_DY, N_ATDY, N_COUNT
};
struct CheckF {
  static constexpr bool kWriteW = false, kWriteE = false;
  static constexpr int kPlaneOut = 3;  // A'y of the plane rows (delta_y is only kept for agents without planes)
  __device__ __forceinline__ void plane_begin(const double (&in)[3]) {
    xv[VX] = in[0]; xv[VY] = in[1]; xv[VP] = in[2];
    acc[VX] = 0.0; acc[VY] = 0.0; acc[VP] = 0.0;
#pragma unroll
    for (int k = 0; k < N_COUNT; ++k) nrm[k] = 0.0;
    ineq_lhs = 0.0;
  }
  __device__ __forceinline__ void plane_emit(double *o) const { o[0] = acc[VX]; o[1] = acc[VY]; o[2] = acc[VP]; }
  __device__ __forceinline__ void plane_absorb(const double *o) { acc[VX] += o[0]; acc[VY] += o[1]; acc[VP] += o[2]; }
  __device__ __forceinline__ void plane_merge(const CheckF &g) {  // row norms: maxima, any order
#pragma unroll
    for (int k = 0; k < N_COUNT; ++k) nrm[k] = fmax(nrm[k], g.nrm[k]);
  }
  double xv[10];
  double acc[10];   // A'y (raw-space accumulation)
  double accd[10];  // A'delta_y
  double nrm[N_COUNT];
  double ineq_lhs;
  double rho;
  bool with_dy;
  const double *dy_base;
  int dy_stride;
  template <int NC, bool EQ = false>
  __device__ __forceinline__ void row(int rid, int i0, double c0, int i1, double c1, int i2, double c2, int i3,
                                      double c3, double l, double u, double &w, double &E) {
    const double e = E, einv = 1.0 / e;
    const double ls = e * l, us = e * u;
    const double rho_i = row_rho(ls, us, rho);
    const double ax = e * row_dot<NC>(xv, i0, c0, i1, c1, i2, c2, i3, c3);
    const double wv = w, z = clipd(wv, ls, us);
    const double y = rho_i * (wv - z);
    const double r = ax - z;
    nrm[N_PRI_S] = fmax(nrm[N_PRI_S], fabs(r));
    nrm[N_Z_S] = fmax(nrm[N_Z_S], fabs(z));
    nrm[N_AX_S] = fmax(nrm[N_AX_S], fabs(ax));
    nrm[N_PRI_U] = fmax(nrm[N_PRI_U], fabs(einv * r));
    nrm[N_Z_U] = fmax(nrm[N_Z_U], fabs(einv * z));
    nrm[N_AX_U] = fmax(nrm[N_AX_U], fabs(einv * ax));
    row_scatter<NC>(acc, e * y, i0, c0, i1, c1, i2, c2, i3, c3);
    if (with_dy) {
      const double dy = rid >= 0 ? dy_base[rid * dy_stride] : 0.0;
      nrm[N_DY] = fmax(nrm[N_DY], fabs(e * dy));
      ineq_lhs += us * (dy > 0 ? dy : 0) + ls * (dy < 0 ? dy : 0);
      row_scatter<NC>(accd, e * dy, i0, c0, i1, c1, i2, c2, i3, c3);
    }
  }
};

// one Ruiz pass over the rows (OSQP scale_data): row norms -> E, column maxima
struct ScaleF {
  static constexpr bool kWriteW = false, kWriteE = true;
  static constexpr int kPlaneOut = 3;  // column maxima on x, y, yaw of the plane's step
  __device__ __forceinline__ void plane_begin(const double (&in)[3]) {
    dv[VX] = in[0]; dv[VY] = in[1]; dv[VP] = in[2];
    cmax[VX] = 0.0; cmax[VY] = 0.0; cmax[VP] = 0.0;
  }
  __device__ __forceinline__ void plane_emit(double *o) const { o[0] = cmax[VX]; o[1] = cmax[VY]; o[2] = cmax[VP]; }
  __device__ __forceinline__ void plane_absorb(const double *o) {
    cmax[VX] = fmax(cmax[VX], o[0]); cmax[VY] = fmax(cmax[VY], o[1]); cmax[VP] = fmax(cmax[VP], o[2]);
  }
  __device__ __forceinline__ void plane_merge(const ScaleF &) {}
  double dv[10];    // current D of the touched unknowns
  double cmax[10];  // max_i E_i |a_ij| per touched unknown (without D_j)
  template <int NC, bool EQ = false>
  __device__ __forceinline__ void row(int rid, int i0, double c0, int i1, double c1, int i2, double c2, int i3,
                                      double c3, double l, double u, double &w, double &E) {
    const double e = E;
    double rn = fabs(c0) * dv[i0];
    cmax[i0] = fmax(cmax[i0], e * fabs(c0));
    if (NC > 1) { rn = fmax(rn, fabs(c1) * dv[i1]); cmax[i1] = fmax(cmax[i1], e * fabs(c1)); }
    if (NC > 2) { rn = fmax(rn, fabs(c2) * dv[i2]); cmax[i2] = fmax(cmax[i2], e * fabs(c2)); }
    if (NC > 3) { rn = fmax(rn, fabs(c3) * dv[i3]); cmax[i3] = fmax(cmax[i3], e * fabs(c3)); }
    rn = limit_scaling(e * rn);
    E = e * (1.0 / sqrt(rn));
  }
};

// reset E to 1 before scaling
struct ResetF {
  static constexpr bool kWriteW = true, kWriteE = true;
  static constexpr int kPlaneOut = 0;
  __device__ __forceinline__ void plane_begin(const double (&)[3]) {}
  __device__ __forceinline__ void plane_emit(double *) const {}
  __device__ __forceinline__ void plane_absorb(const double *) {}
  __device__ __forceinline__ void plane_merge(const ResetF &) {}
  template <int NC, bool EQ = false>
  __device__ __forceinline__ void row(int, int, double, int, double, int, double, int, double, double, double,
                                      double &w, double &E) {
    E = 1.0;
    w = 0.0;
  }
};

// A' diag(rho E^2) A in raw space: own 6x6 block (lower triangle), the coupling
// to the next step and the next step's diagonal contributions.
struct HasmF {
  static constexpr bool kWriteW = false, kWriteE = false;
  static constexpr int kPlaneOut = 6;  // lower triangle of the (x, y, yaw) block
  __device__ __forceinline__ void plane_begin(const double (&)[3]) {
    q[VX][VX] = 0.0; q[VY][VX] = 0.0; q[VY][VY] = 0.0; q[VP][VX] = 0.0; q[VP][VY] = 0.0; q[VP][VP] = 0.0;
  }
  __device__ __forceinline__ void plane_emit(double *o) const {
    o[0] = q[VX][VX]; o[1] = q[VY][VX]; o[2] = q[VY][VY]; o[3] = q[VP][VX]; o[4] = q[VP][VY]; o[5] = q[VP][VP];
  }
  __device__ __forceinline__ void plane_absorb(const double *o) {
    q[VX][VX] += o[0]; q[VY][VX] += o[1]; q[VY][VY] += o[2]; q[VP][VX] += o[3]; q[VP][VY] += o[4]; q[VP][VP] += o[5];
  }
  __device__ __forceinline__ void plane_merge(const HasmF &) {}
  double q[6][6];   // q[i][j], j <= i
  double cr[4][6];  // rows = next-step x,y,yaw,steer; cols = own unknowns
  double nd[4];
  double rho;
  __device__ __forceinline__ void add(int i, int j, double v) {
    if (i < j) { int s = i; i = j; j = s; }
    if (i < 6) q[i][j] += v;
    else if (j < 6) cr[i - 6][j] += v;
    else nd[i - 6] += v;  // only i == j occurs
  }
  template <int NC, bool EQ = false>
  __device__ __forceinline__ void row(int rid, int i0, double c0, int i1, double c1, int i2, double c2, int i3,
                                      double c3, double l, double u, double &w, double &E) {
    const double e = E;
    const double h = row_rho(e * l, e * u, rho) * (e * e);
    add(i0, i0, h * c0 * c0);
    if (NC > 1) { add(i1, i0, h * c1 * c0); add(i1, i1, h * c1 * c1); }
    if (NC > 2) { add(i2, i0, h * c2 * c0); add(i2, i1, h * c2 * c1); add(i2, i2, h * c2 * c2); }
    if (NC > 3) { add(i3, i0, h * c3 * c0); add(i3, i1, h * c3 * c1); add(i3, i2, h * c3 * c2); add(i3, i3, h * c3 * c3); }
  }
};

// ===================================================================
// QP phases
// ===================================================================
__device__ __forceinline__ void load_xv(const Ctx &c, const double *v, double (&xv)[10]) {
  const int NT = c.NT(), t = c.t();
#pragma unroll
  for (int k = 0; k < 6; ++k) xv[k] = v[k * NT + t];
#pragma unroll
  for (int k = 0; k < 4; ++k) xv[6 + k] = c.has_next() ? v[k * NT + t + 1] : 0.0;
}

// number of unknowns of step t (the last step has no v, w)
__device__ __forceinline__ int nvar(const Ctx &c) { return c.has_next() ? 6 : 4; }

// linearization-dependent per-step data (dsqp_solver.cc:670-718, 893-948, 1116-1123)
__device__ __forceinline__ void assemble_rows(Ctx &c, const csdo_params &P) {
  const int NT = c.NT(), Nt = c.Nt(), t = c.t();
  if (c.active()) {
    const double yaw0 = c.cur()[2 * NT + t], st0 = c.cur()[3 * NT + t], v0 = c.cur()[4 * NT + t];
    const double sn = sin(yaw0), cs = cos(yaw0);
    c.ros()[RO_SN * NT + c.t()] = sn;
    c.ros()[RO_CS * NT + c.t()] = cs;
    if (c.has_next()) {
      const double cd = cos(st0);
      c.ros()[RO_A1 * NT + c.t()] = -P.dt * (v0 * sn);
      c.ros()[RO_A2 * NT + c.t()] = P.dt * (v0 * cs);
      c.ros()[RO_A3 * NT + c.t()] = (P.dt / P.WB * v0) / (cd * cd);
      c.ros()[RO_B3 * NT + c.t()] = P.dt / P.WB * tan(st0);
      c.ros()[RO_KR0 * NT + c.t()] = -(P.dt * yaw0 * v0 * sn);
      c.ros()[RO_KR1 * NT + c.t()] = -(-P.dt * yaw0 * v0 * cs);
      c.ros()[RO_KR2 * NT + c.t()] = -(-P.dt * (st0 * v0 / P.WB / (cd * cd)));
    }
    const double dxf = -P.f2x * sn, dyf = P.f2x * cs, dxr = -P.r2x * sn, dyr = P.r2x * cs;
    const double exf = P.f2x * (cs + yaw0 * sn), eyf = P.f2x * (sn - yaw0 * cs);
    const double exr = P.r2x * (cs + yaw0 * sn), eyr = P.r2x * (sn - yaw0 * cs);
    c.ros()[RO_CL0 * NT + c.t()] = __ldcg(&c.corr()[0 * Nt + t]) - exf; c.ros()[RO_CU0 * NT + c.t()] = __ldcg(&c.corr()[1 * Nt + t]) - exf;
    c.ros()[RO_CL1 * NT + c.t()] = __ldcg(&c.corr()[2 * Nt + t]) - eyf; c.ros()[RO_CU1 * NT + c.t()] = __ldcg(&c.corr()[3 * Nt + t]) - eyf;
    c.ros()[RO_CL2 * NT + c.t()] = __ldcg(&c.corr()[4 * Nt + t]) - exr; c.ros()[RO_CU2 * NT + c.t()] = __ldcg(&c.corr()[5 * Nt + t]) - exr;
    c.ros()[RO_CL3 * NT + c.t()] = __ldcg(&c.corr()[6 * Nt + t]) - eyr; c.ros()[RO_CU3 * NT + c.t()] = __ldcg(&c.corr()[7 * Nt + t]) - eyr;
    for (int k = c.pstart()[t]; k < c.pstart()[t + 1]; ++k) {
      const double *pl = c.plane_abc() + (size_t)12 * k;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const double a = pl[3 * r], b = pl[3 * r + 1], cc = pl[3 * r + 2];
        const double dx = r < 2 ? dxf : dxr, dy = r < 2 ? dyf : dyr;
        const double ex = r < 2 ? exf : exr, ey = r < 2 ? eyf : eyr;
        double *q = c.pl() + (size_t)PL_COUNT * (4 * k + r);
        q[PL_A] = a * 1.0;
        q[PL_B] = b * 1.0;
        q[PL_G] = a * dx + b * dy;
        q[PL_U] = -(cc + (a * ex + b * ey));
      }
    }
  }
  __syncthreads();
}

// OSQP scale_data, `scaling` Ruiz passes; leaves D, E, c
__device__ __forceinline__ void ruiz_scale(Ctx &c, const csdo_params &P) {
  const int NT = c.NT(), t = c.t(), Nt = c.Nt();
  ResetF rf;
  if (c.active()) {
    visit_rows(c, P, rf);
#pragma unroll
    for (int k = 0; k < 6; ++k) c.D()[k * NT + t] = 1.0;
  }
  visit_planes<IN_NONE>(c, P, rf);
  c.c = 1.0;
  __syncthreads();
  for (int pass = 0; pass < P.scaling; ++pass) {
    double dt_new[6];
    ScaleF sf;
    if (c.active()) {
      load_xv(c, c.D(), sf.dv);
#pragma unroll
      for (int k = 0; k < 10; ++k) sf.cmax[k] = 0.0;
      visit_rows(c, P, sf);
    }
    visit_planes<IN_D>(c, P, sf);
    if (c.active()) {
#pragma unroll
      for (int k = 0; k < 4; ++k) c.carry()[k * NT + t] = sf.cmax[6 + k];
#pragma unroll
      for (int k = 0; k < 6; ++k) dt_new[k] = sf.cmax[k];
    }
    __syncthreads();
    if (c.active()) {
      const int nv = nvar(c);
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        if (k >= nv) continue;
        const double dk = c.D()[k * NT + t];
        double a = dt_new[k];
        if (k < 4 && t > 0) a = fmax(a, c.carry()[k * NT + t - 1]);
        double colA = a * dk;  // inf-norm of column k of the scaled A
        double colP = 0.0;     // inf-norm of the column of the scaled (symmetric) P
        if (k == VV) {
          const double pv = (t != 0 && t != Nt - 2) ? 2.0 : 1.0;
          colP = pv * dk;
          if (t > 0) colP = fmax(colP, c.D()[VV * NT + t - 1]);
          if (t < Nt - 2) colP = fmax(colP, c.D()[VV * NT + t + 1]);
          colP = c.c * dk * colP;
        } else if (k == VW) {
          colP = c.c * dk * dk;
        }
        dt_new[k] = 1.0 / sqrt(limit_scaling(fmax(colP, colA)));
      }
    }
    __syncthreads();
    if (c.active()) {
      const int nv = nvar(c);
#pragma unroll
      for (int k = 0; k < 6; ++k)
        if (k < nv) c.D()[k * NT + t] *= dt_new[k];
    }
    __syncthreads();
    // cost normalisation: mean column norm of the new P (q = 0 -> its norm counts as 1)
    double s[1] = {0.0};
    if (c.active() && c.has_next()) {
      const double dvv = c.D()[VV * NT + t], dw = c.D()[VW * NT + t];
      const double pv = (t != 0 && t != Nt - 2) ? 2.0 : 1.0;
      double colP = pv * dvv;
      if (t > 0) colP = fmax(colP, c.D()[VV * NT + t - 1]);
      if (t < Nt - 2) colP = fmax(colP, c.D()[VV * NT + t + 1]);
      s[0] = c.c * dvv * colP + c.c * dw * dw;
    }
    block_reduce<1, false>(s, c.red());
    double c_temp = s[0] / (double)(6 * Nt - 2);
    const double inf_norm_q = 1.0;  // limit_scaling(0) == 1
    c_temp = fmax(c_temp, inf_norm_q);
    c_temp = limit_scaling(c_temp);
    c.c *= 1.0 / c_temp;
  }
}

// reduced KKT  H = c D P D + sigma I + D A_raw' diag(rho E^2) A_raw D  into the
// band storage, then LDL'.
// block t's row k of the band storage: rec[d], d = 1..6, is H_{i,i-d} (i = 6t + k), diag its H_ii slot.
// OW: the one-warp solver's row storage (band_solver.cuh), else the block records of pbcr_solver.cuh.
using OwSolveFn = void (*)(const BandMem, double *, double *, int, int);
using OwFactorFn = void (*)(const BandMem, int);
// register class RC of a kernel variant: 0 = 255 registers, 1 = 168, 2 = 128 with the CTA-wide solver
// (pbcr_solver.cuh); 3 = 255, 4 = 168 with the one-warp solver (band_solver.cuh, block sizes <= 96)
__host__ __device__ constexpr bool one_warp(int RC) { return RC >= 3; }
#ifdef CSDO_OW_ROWS_COLD
constexpr bool kOwRowsInline = false;
#else
constexpr bool kOwRowsInline = true;
#endif

template <bool OW>
struct BandRow {
  double *rec, *diag;
  __device__ __forceinline__ BandRow(const Ctx &c, int t, int k) {
    if (OW) {
      const int sk = make_parts(c.Nt(), c.s->skew_tab).skew_of_block(t);
      rec = c.bm().L6 + sk + (size_t)(6 * t + k) * 6 - 1;
      diag = c.bm().dinv + 6 * t + k;
    } else {
      double *B = pbcr_blk(c.pm().L, t);
      rec = B + 6 * k - 1;
      diag = B + 36 + k;
    }
  }
};

template <bool OW>
__device__ __forceinline__ void form_and_factor(Ctx &c, const csdo_params &P) {
  const int NT = c.NT(), t = c.t(), Nt = c.Nt();
  HasmF hf;
  hf.rho = c.rho;
  if (c.active()) {
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = 0; j < 6; ++j) hf.q[i][j] = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      hf.nd[i] = 0.0;
#pragma unroll
      for (int j = 0; j < 6; ++j) hf.cr[i][j] = 0.0;
    }
    visit_rows(c, P, hf);
  }
  visit_planes<IN_NONE>(c, P, hf);
  if (c.active()) {
#pragma unroll
    for (int k = 0; k < 4; ++k) c.carry()[k * NT + t] = hf.nd[k];
    // clear this step's block record; the last step's missing v, w are dummy unknowns (H_ii = 1)
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const BandRow<OW> r(c, t, k);
#pragma unroll
      for (int d = 1; d <= 6; ++d) r.rec[d] = 0.0;
      *r.diag = 1.0;
    }
  }
  __syncthreads();
  if (c.active()) {
    const int nv = nvar(c);
    double dk[10];
    load_xv(c, c.D(), dk);
    // objective (dsqp_solver.cc:163-197): second difference on v, identity on w
    if (c.has_next()) {
      hf.q[VV][VV] += c.c * ((t != 0 && t != Nt - 2) ? 2.0 : 1.0);
      hf.q[VW][VW] += c.c * 1.0;
    }
    for (int k = 0; k < nv; ++k) {
      const BandRow<OW> r(c, t, k);  // r.rec[d] = H_{i,i-d}
      double diag = hf.q[k][k];
      if (k < 4 && t > 0) diag += c.carry()[k * NT + t - 1];
      *r.diag = dk[k] * dk[k] * diag + P.sigma;
      for (int j = 0; j < k; ++j) r.rec[k - j] = dk[k] * dk[j] * hf.q[k][j];
    }
    if (c.has_next()) {
      const int nvn = (t + 1 < Nt - 1) ? 6 : 4;
      for (int k = 0; k < 4; ++k) {  // rows x,y,yaw,steer of step t+1
        const BandRow<OW> r(c, t + 1, k);
        for (int j = k; j < 6; ++j) r.rec[6 + k - j] = dk[6 + k] * dk[j] * hf.cr[k][j];
      }
      if (nvn == 6)  // v_{t+1} - v_t coupling of the objective
        BandRow<OW>(c, t + 1, VV).rec[6] = c.D()[VV * NT + t + 1] * dk[VV] * (-c.c);
    }
  }
  __syncthreads();
  if (OW) {
    // one warp factors (the others wait): entered through a pointer so that it gets its own register budget
    __syncwarp();
#ifdef CSDO_OW_DIRECT
    if ((c.tid() >> 5) == c.s->solver_warp) {
      if (c.l_shared()) band_factor_warp<true>(c.bm(), Nt);
      else band_factor_warp<false>(c.bm(), Nt);
    }
#else
    if ((c.tid() >> 5) == c.s->solver_warp) reinterpret_cast<OwFactorFn>(c.s->fn_factor)(c.bm(), Nt);
#endif
    __syncthreads();
  } else {
    if (c.l_shared()) pbcr_factor_cta<true>(c.pm(), Nt);
    else pbcr_factor_cta<false>(c.pm(), Nt);
  }
}

// optional phase timing (thread 0's clock), accumulated per CTA and added to queue[2..] at exit
#ifdef CSDO_DEV_TIMERS
#define PH_T0() long long ph_t0 = clock64()
#define PH_ADD(id) do { const long long ph_t1 = clock64(); c.ph[id] += ph_t1 - ph_t0; ph_t0 = ph_t1; } while (0)
#define PH_RESET() ph_t0 = clock64()
#else
#define PH_T0() do {} while (0)
#define PH_ADD(id) do {} while (0)
#define PH_RESET() do {} while (0)
#endif

struct QpOut {
  int status, iters, n_factor;
};

// rhs <- sigma x + D (acc + carry[t-1])   (q = 0)
__device__ __forceinline__ void finish_rhs(const Ctx &c, const csdo_params &P, const double (&acc)[10]) {
  const int NT = c.NT(), t = c.t();
  if (c.active()) {
    const int nv = nvar(c);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      if (k >= nv) { c.rhs()[k * NT + t] = 0.0; continue; }  // dummy unknowns stay 0
      double a = acc[k];
      if (k < 4 && t > 0) a += c.carry()[k * NT + t - 1];
      c.rhs()[k * NT + t] = P.sigma * c.x()[k * NT + t] + c.D()[k * NT + t] * a;
    }
  }
}

// AFTER_SOLVE: x~ sits in rhs (scaled space).  The row pass then forms D x~ of this and the next step
// itself and does the relaxed x update of its own step on the way (no separate pass, no barrier).
// RR (optional): the step's row data / scaling / state held in registers by the caller across ADMM
// iterations (then nothing of the fixed rows is read from or written to memory here)
template <int MODE, bool AFTER_SOLVE = false>
__device__ __forceinline__ void step_rows(Ctx &c, const csdo_params &P, bool store_dy, double rho_old,
                                          RowRegs *RR = nullptr) {
  DBG_INIT();
  StepF<MODE> sf;
  sf.alpha = P.alpha; sf.rho = c.rho; sf.rho_old = rho_old;
  sf.store_dy = store_dy; sf.dy_base = c.dy() + c.t(); sf.dy_stride = c.NT();
  if (c.active()) {
    if (AFTER_SOLVE) {
      const int NT = c.NT(), t = c.t(), nv = nvar(c);
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const double xtil = c.rhs()[k * NT + t];
        sf.xv[k] = c.D()[k * NT + t] * xtil;
        if (k < nv) c.x()[k * NT + t] = P.alpha * xtil + (1.0 - P.alpha) * c.x()[k * NT + t];
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) sf.xv[6 + k] = c.has_next() ? c.D()[k * NT + t + 1] * c.rhs()[k * NT + t + 1] : 0.0;
    } else {
      load_xv(c, c.xt(), sf.xv);
    }
#pragma unroll
    for (int k = 0; k < 10; ++k) sf.acc[k] = 0.0;
    if (RR) rows_apply(c, P, sf, *RR);
    else visit_rows(c, P, sf);
  }
  if (MODE == 2) DBG_ACC(8);    // fixed rows (thread 0's own)
  visit_planes<MODE == 3 ? IN_NONE : (AFTER_SOLVE ? IN_DXT : IN_XT)>(c, P, sf);
  if (MODE == 2) DBG_ACC(9);    // plane-major pass + barrier + absorb
  if (c.active()) {
#pragma unroll
    for (int k = 0; k < 4; ++k) c.carry()[k * c.NT() + c.t()] = sf.acc[6 + k];
  }
  __syncthreads();  // every thread has read its neighbour's x~ before rhs is rebuilt
  if (MODE == 2) DBG_ACC(10);   // barrier: waiting for the slowest thread of the row pass
  finish_rhs(c, P, sf.acc);
  __syncthreads();
  if (MODE == 2) DBG_ACC(11);   // finish_rhs + barrier
}

// OSQP update_info + check_termination(approximate = false/true) on the reduced scalars
struct CheckOut {
  double nrm[N_COUNT];
  double ineq_lhs;
};

// (a __noinline__ copy shrinks the code but shifts the register allocation of the row pass: no net gain)
__device__ __forceinline__ void check_rows(Ctx &c, const csdo_params &P, bool with_dy, CheckOut &co) {
  const int NT = c.NT(), t = c.t(), Nt = c.Nt();
  // xt <- D x (the current iterate, not x~)
  if (c.active()) {
#pragma unroll
    for (int k = 0; k < 6; ++k) c.xt()[k * NT + t] = c.D()[k * NT + t] * c.x()[k * NT + t];
  }
  __syncthreads();
  CheckF cf;
#pragma unroll
  for (int k = 0; k < N_COUNT; ++k) cf.nrm[k] = 0.0;
  cf.ineq_lhs = 0.0;
  cf.rho = c.rho; cf.with_dy = with_dy; cf.dy_base = c.dy() + c.t(); cf.dy_stride = c.NT();
  if (c.active()) {
    load_xv(c, c.xt(), cf.xv);
#pragma unroll
    for (int k = 0; k < 10; ++k) { cf.acc[k] = 0.0; cf.accd[k] = 0.0; }
    visit_rows(c, P, cf);
  }
  visit_planes<IN_XT>(c, P, cf);
  if (c.active()) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      c.carry()[k * NT + t] = cf.acc[6 + k];
      c.rhs()[k * NT + t] = cf.accd[6 + k];  // rhs is rebuilt below, see caller
    }
  }
  __syncthreads();
  if (c.active()) {
    const int nv = nvar(c);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      if (k >= nv) continue;
      const double dk = c.D()[k * NT + t], dinv = 1.0 / dk;
      double a = cf.acc[k], ad = cf.accd[k];
      if (k < 4 && t > 0) { a += c.carry()[k * NT + t - 1]; ad += c.rhs()[k * NT + t - 1]; }
      const double aty = dk * a;
      double px = 0.0;
      if (k == VV) {
        const double pv = (t != 0 && t != Nt - 2) ? 2.0 : 1.0;
        double sacc = pv * (dk * c.x()[VV * NT + t]);
        if (t > 0) sacc -= c.D()[VV * NT + t - 1] * c.x()[VV * NT + t - 1];
        if (t < Nt - 2) sacc -= c.D()[VV * NT + t + 1] * c.x()[VV * NT + t + 1];
        px = c.c * dk * sacc;
      } else if (k == VW) {
        px = c.c * dk * (dk * c.x()[VW * NT + t]);
      }
      const double r = px + aty;
      cf.nrm[N_DUA_S] = fmax(cf.nrm[N_DUA_S], fabs(r));
      cf.nrm[N_PX_S] = fmax(cf.nrm[N_PX_S], fabs(px));
      cf.nrm[N_ATY_S] = fmax(cf.nrm[N_ATY_S], fabs(aty));
      cf.nrm[N_DUA_U] = fmax(cf.nrm[N_DUA_U], fabs(dinv * r));
      cf.nrm[N_PX_U] = fmax(cf.nrm[N_PX_U], fabs(dinv * px));
      cf.nrm[N_ATY_U] = fmax(cf.nrm[N_ATY_U], fabs(dinv * aty));
      if (with_dy) cf.nrm[N_ATDY] = fmax(cf.nrm[N_ATDY], fabs(dinv * (dk * ad)));
    }
  }
  __syncthreads();
  block_reduce<N_COUNT, true>(cf.nrm, c.red());
  double s[1] = {cf.ineq_lhs};
  block_reduce<1, false>(s, c.red());
#pragma unroll
  for (int k = 0; k < N_COUNT; ++k) co.nrm[k] = cf.nrm[k];
  co.ineq_lhs = s[0];
}

// check_termination (OSQP auxil.c); returns status or 0 when not terminated
static __device__ int termination_status(const Ctx &c, const csdo_params &P, const CheckOut &co, bool approximate) {
  const double cinv = 1.0 / c.c;
  double eps_abs = P.eps_abs, eps_rel = P.eps_rel, eps_pinf = P.eps_prim_inf;
  const double pri_res = co.nrm[N_PRI_U], dua_res = cinv * co.nrm[N_DUA_U];
  if (pri_res > kOsqpInfty || dua_res > kOsqpInfty) return CSDO_QP_NON_CVX;
  if (approximate) { eps_abs *= 10; eps_rel *= 10; eps_pinf *= 10; }
  const double eps_prim = eps_abs + eps_rel * fmax(co.nrm[N_Z_U], co.nrm[N_AX_U]);
  bool prim_ok = false, prim_inf = false;
  if (pri_res < eps_prim) prim_ok = true;
  else if (c.K() == 0) {
    // is_primal_infeasible; with K > 0 every inter-vehicle row carries a true
    // -inf lower bound and the support-function sum is NaN, so the test never fires
    const double ndy = co.nrm[N_DY];
    if (ndy > eps_pinf && co.ineq_lhs < -eps_pinf * ndy) prim_inf = co.nrm[N_ATDY] < eps_pinf * ndy;
  }
  const double eps_dual = eps_abs + eps_rel * (cinv * fmax(co.nrm[N_ATY_U], co.nrm[N_PX_U]));
  const bool dual_ok = dua_res < eps_dual;  // q = 0: dual infeasibility cannot trigger
  if (prim_ok && dual_ok) return approximate ? CSDO_QP_SOLVED_INACCURATE : CSDO_QP_SOLVED;
  if (prim_inf) return approximate ? CSDO_QP_PRIMAL_INFEASIBLE_INACCURATE : CSDO_QP_PRIMAL_INFEASIBLE;
  return 0;
}

// -------------------------------------------------------------------------------------------------
// Out-of-line ("cold") entry points.  Everything that runs once per QP, once per 25 ADMM iterations or
// once per SQP iteration is kept OUT of the ADMM loop's function body: inlined, those phases (the
// factorization alone is 12 k SASS instructions, and it was inlined twice) set the register allocation of
// the whole kernel, and the hot loop (band solve + row pass) ended up re-loading spilled scalars from
// local memory at L2 latency.  The wrappers take the shared context pointer and the two per-QP scalars by
// value (a by-reference Ctx would live in local memory) and read the parameters from the context.
__device__ __forceinline__ Ctx make_ctx(CtxShared *s, double rho, double cc) {
  __builtin_assume(__isShared(s));
  Ctx c;
  c.s = s; c.rho = rho; c.c = cc;
#ifdef CSDO_DEV_TIMERS
  for (int k = 0; k < 8; ++k) c.ph[k] = 0;
#endif
  return c;
}
template <int RC>
__device__ __noinline__ double ruiz_scale_cold(CtxShared *s) {
  Ctx c = make_ctx(s, 0.0, 1.0);
  ruiz_scale(c, s->P);
  return c.c;
}
template <int RC>
__device__ __noinline__ void form_and_factor_cold(CtxShared *s, double rho, double cc) {
  Ctx c = make_ctx(s, rho, cc);
  form_and_factor<one_warp(RC)>(c, s->P);
}
template <int RC>
__device__ __noinline__ void assemble_rows_cold(CtxShared *s) {
  Ctx c = make_ctx(s, 0.0, 1.0);
  assemble_rows(c, s->P);
}
template <int RC, int MODE, bool AFTER_SOLVE>
__device__ __noinline__ void step_rows_cold(CtxShared *s, double rho, double cc, bool store_dy, double rho_old) {
  Ctx c = make_ctx(s, rho, cc);
  step_rows<MODE, AFTER_SOLVE>(c, s->P, store_dy, rho_old);
}
template <int RC>
__device__ __noinline__ void check_rows_cold(CtxShared *s, double rho, double cc, bool with_dy, CheckOut *co) {
  Ctx c = make_ctx(s, rho, cc);
  check_rows(c, s->P, with_dy, *co);
}
template <int RC>
__device__ __noinline__ int termination_status_cold(CtxShared *s, double cc, const CheckOut *co, bool approximate) {
  Ctx c = make_ctx(s, 0.0, cc);
  return termination_status(c, s->P, *co, approximate);
}

// The band solve as its own function.  Its sweeps want ~200 registers; a noinline function called directly is
// capped by what is live in its caller (1.1 KB of spills inside the sweeps), while a call through a function
// pointer read from device memory follows the full ABI and gives the callee the whole register file.  Which
// part goes behind the pointer depends on the register class, see the body.  Tried as well: the sweeps behind
// the pointer and the rest of the solve inline in the ADMM loop -- 2-4 % slower on every workload.
template <int RC>
__device__ __noinline__ void band_solve_call(CtxShared *s) {
  __builtin_assume(__isShared(s));
  const PbcrMem pm = s->pm;
  if (RC == 0) {
    // 255-register variants: this function is called DIRECTLY and stays light (separator phases only); the two
    // 3-block sweeps, which want ~200 registers, go through a pointer and only the Nt/4 threads that own a
    // partition pay the callee-saved save/restore (measured +5 % at Nt >= 128 against the whole solve behind
    // the pointer, where all 256 threads push 110 KB of registers through L2 per solve)
    if (s->l_shared) pbcr_solve_cta<true, true>(pm, s->rhs, s->xt, s->Nt, s->NT, s->fn_sweep);
    else pbcr_solve_cta<false, true>(pm, s->rhs, s->xt, s->Nt, s->NT, s->fn_sweep);
  } else {
    // 168- / 128-register variants: the whole solve behind the pointer (measured 5 % faster at Nt <= 96, 3 CTAs/SM)
    if (s->l_shared) pbcr_solve_cta<true>(pm, s->rhs, s->xt, s->Nt, s->NT);
    else pbcr_solve_cta<false>(pm, s->rhs, s->xt, s->Nt, s->NT);
  }
}
using BandSolveFn = void (*)(CtxShared *);
// The function-pointer tables are per translation unit ON PURPOSE (see the note at the top of this file): with
// the one-warp solver's entry points address-taken in the same module, ptxas gave every out-of-line phase of
// the 255-register kernels a larger save area (step_rows_cold: 64 -> 184 B of spills) and Nt = 256 lost 4 %.
#if CSDO_TU == 2
// (band_solve_call<0> is called directly but stays in the table: with its address not taken -- a third module
// for the 168- / 128-register variants was tried -- it needs no register saves itself and its callers save
// more instead, 1.5-2.7 % slower at Nt = 128..256)
__device__ BandSolveFn g_band_solve[3] = {band_solve_call<0>, band_solve_call<1>, band_solve_call<2>};
__device__ PbcrSweepFn g_pbcr_sweep[2] = {pbcr_sweep_entry<false>, pbcr_sweep_entry<true>};
#endif
#if CSDO_TU == 1 && !defined(CSDO_OW_DIRECT)
__device__ OwSolveFn g_ow_solve[2] = {band_solve_warp<false>, band_solve_warp<true>};
__device__ OwFactorFn g_ow_factor[2] = {band_factor_warp<false>, band_factor_warp<true>};
#endif

// solveOSQP (dsqp_solver.cc:423-555): setup + warm start + ADMM; solution in c.sol()


// RC: register class of the kernel variant (0: 255, 1: 168, 2: 128 registers).  The out-of-line phases are
// instantiated per class: a function shared by all variants would be compiled for the smallest budget.
template <int RC>
__device__ __forceinline__ QpOut solve_qp(Ctx &c, const csdo_params &P, const Layout &LY) {
  const int NT = c.NT(), t = c.t(), Nt = c.Nt();
  QpOut out{CSDO_QP_UNSOLVED, 0, 1};
  PH_T0();
  c.c = ruiz_scale_cold<RC>(c.s);
  PH_ADD(2);
  c.rho = fmin(fmax(P.rho, kRhoMin), kRhoMax);
  form_and_factor_cold<RC>(c.s, c.rho, c.c);
  PH_ADD(3);
  // osqp_warm_start_x: x <- Dinv x0, z <- A x, y = 0
  if (c.active()) {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const double dk = c.D()[k * NT + t];
      const double xs = c.cur()[k * NT + t] * (1.0 / dk);
      c.x()[k * NT + t] = xs;
      c.xt()[k * NT + t] = dk * xs;
    }
  }
  __syncthreads();
  step_rows_cold<RC, 0, false>(c.s, c.rho, c.c, false, 0.0);
  const bool keep_dy = (c.K() == 0);
  CheckOut co;
  bool checked = false;
  int iter = 0;
  PH_ADD(5);
  for (iter = 1; iter <= P.osqp_max_iter; ++iter) {
    if constexpr (one_warp(RC)) {
      __syncwarp();  // the solver warp must enter the solve converged (threads leave barriers individually)
#ifdef CSDO_OW_DIRECT   // (developer switch: direct calls instead of the function pointers)
      if ((c.tid() >> 5) == c.s->solver_warp) {
        if (c.l_shared()) band_solve_warp<true>(c.s->bm, c.rhs(), c.xt(), Nt, NT);
        else band_solve_warp<false>(c.s->bm, c.rhs(), c.xt(), Nt, NT);
      }
#else
      if ((c.tid() >> 5) == c.s->solver_warp)
        reinterpret_cast<OwSolveFn>(c.s->fn_solve)(c.s->bm, c.rhs(), c.xt(), Nt, NT);
#endif
#ifdef CSDO_DOUBLE_SOLVE  // timing experiment: a second, discarded solve (its cost is the solve's share of the step)
      if ((c.tid() >> 5) == c.s->solver_warp)
        reinterpret_cast<OwSolveFn>(c.s->fn_solve)(c.s->bm, c.xt(), c.xt(), Nt, NT);
#endif
      __syncthreads();
    } else if constexpr (RC == 0) {
      band_solve_call<0>(c.s);
    } else {
      reinterpret_cast<BandSolveFn>(c.s->fn_solve)(c.s);
    }
    PH_ADD(4);
    const bool can_check = P.check_termination && (iter % P.check_termination == 0);
    const bool store_dy = keep_dy && (can_check || iter == P.osqp_max_iter);
    if (iter == 1) step_rows_cold<RC, 1, true>(c.s, c.rho, c.c, store_dy, 0.0);
#ifdef CSDO_ROWS_INLINE   // (developer switch: the hot row pass inline measured 2-3 % slower than out of line)
    else step_rows<2, true>(c, P, store_dy, 0.0);
#else
    else if (one_warp(RC) && kOwRowsInline) step_rows<2, true>(c, P, store_dy, 0.0);
    else step_rows_cold<RC, 2, true>(c.s, c.rho, c.c, store_dy, 0.0);
#endif
    PH_ADD(5);
    checked = false;
    const bool adapt = P.adaptive_rho && P.adaptive_rho_interval && (iter % P.adaptive_rho_interval == 0);
    if (can_check || adapt) {
      // check_rows borrows rhs/carry as scratch: save the next right-hand side in xt afterwards
      double keep[6];
      if (c.active())
#pragma unroll
        for (int k = 0; k < 6; ++k) keep[k] = c.rhs()[k * NT + t];
      __syncthreads();
      check_rows_cold<RC>(c.s, c.rho, c.c, keep_dy && store_dy, &co);
      PH_ADD(6);
      if (c.active())
#pragma unroll
        for (int k = 0; k < 6; ++k) c.rhs()[k * NT + t] = keep[k];
      __syncthreads();
      checked = can_check;
      if (can_check) {
        const int st = termination_status_cold<RC>(c.s, c.c, &co, false);
        if (st != 0) { out.status = st; break; }
      }
      if (adapt) {
        // compute_rho_estimate / adapt_rho on the scaled residual norms
        double pri = co.nrm[N_PRI_S], dua = co.nrm[N_DUA_S];
        pri /= (fmax(co.nrm[N_Z_S], co.nrm[N_AX_S]) + 1e-10);
        dua /= (fmax(co.nrm[N_ATY_S], co.nrm[N_PX_S]) + 1e-10);
        double rho_new = c.rho * sqrt(pri / (dua + 1e-10));
        rho_new = fmin(fmax(rho_new, kRhoMin), kRhoMax);
        if (rho_new > c.rho * P.adaptive_rho_tolerance || rho_new < c.rho / P.adaptive_rho_tolerance) {
          const double rho_old = c.rho;
          c.rho = rho_new;
          out.n_factor++;
          form_and_factor_cold<RC>(c.s, c.rho, c.c);
          PH_ADD(3);
          step_rows_cold<RC, 3, false>(c.s, c.rho, c.c, false, rho_old);

          PH_ADD(5);
        }
      }
    }
  }
  if (iter > P.osqp_max_iter) iter = P.osqp_max_iter;
  out.iters = iter;
  if (out.status == CSDO_QP_UNSOLVED) {
    if (!checked) {
      double keep[6];
      if (c.active())
#pragma unroll
        for (int k = 0; k < 6; ++k) keep[k] = c.rhs()[k * NT + t];
      __syncthreads();
      check_rows_cold<RC>(c.s, c.rho, c.c, keep_dy, &co);
      if (c.active())
#pragma unroll
        for (int k = 0; k < 6; ++k) c.rhs()[k * NT + t] = keep[k];
      __syncthreads();
      const int st = termination_status_cold<RC>(c.s, c.c, &co, false);
      if (st != 0) out.status = st;
    }
    if (out.status == CSDO_QP_UNSOLVED) {
      const int st = termination_status_cold<RC>(c.s, c.c, &co, true);
      out.status = st != 0 ? st : CSDO_QP_MAX_ITER_REACHED;
    }
  }
  // store_solution: x_out = D x for statuses that carry a solution, else keep solution0
  const bool has_solution = !(out.status == CSDO_QP_PRIMAL_INFEASIBLE ||
                              out.status == CSDO_QP_PRIMAL_INFEASIBLE_INACCURATE ||
                              out.status == CSDO_QP_NON_CVX);
  if (c.active()) {
    const int nv = nvar(c);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      double v = 0.0;
      if (k < nv) v = (has_solution && abs(out.status) <= 2) ? c.D()[k * NT + t] * c.x()[k * NT + t] : c.cur()[k * NT + t];
      c.sol()[k * NT + t] = v;
    }
  }
  __syncthreads();
  return out;
}

// isFeasible (dsqp_solver.cc:292-420), fully_check = false, on c.sol()
__device__ __forceinline__ bool is_feasible(Ctx &c, const csdo_params &P) {
  const int NT = c.NT(), t = c.t(), Nt = c.Nt();
  double s[1] = {0.0};
  double mx[2] = {0.0, 0.0};
  if (c.active()) {
    const double x0 = c.sol()[0 * NT + t], y0 = c.sol()[1 * NT + t], yaw0 = c.sol()[2 * NT + t];
    const double cs = cos(yaw0), sn = sin(yaw0);
    if (c.has_next()) {
      const double st0 = c.sol()[3 * NT + t], v0 = c.sol()[4 * NT + t], w0 = c.sol()[5 * NT + t];
      double a = x0 + v0 * cs * P.dt - c.sol()[0 * NT + t + 1]; s[0] += a * a;
      a = y0 + v0 * sn * P.dt - c.sol()[1 * NT + t + 1]; s[0] += a * a;
      a = yaw0 + v0 * tan(st0) / P.WB * P.dt - c.sol()[2 * NT + t + 1]; s[0] += a * a;
      a = st0 + w0 * P.dt - c.sol()[3 * NT + t + 1]; s[0] += a * a;
    }
    const double Y[4] = {x0 + P.f2x * cs, y0 + P.f2x * sn, x0 + P.r2x * cs, y0 + P.r2x * sn};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const double lo = __ldcg(&c.corr()[(2 * q) * Nt + t]), hi = __ldcg(&c.corr()[(2 * q + 1) * Nt + t]);
      if (!(lo <= Y[q])) mx[0] = fmax(mx[0], lo - Y[q]);
      if (!(Y[q] <= hi)) mx[0] = fmax(mx[0], Y[q] - hi);
    }
    for (int k = c.pstart()[t]; k < c.pstart()[t + 1]; ++k) {
      const double *pl = c.plane_abc() + (size_t)12 * k;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const double res = r < 2 ? pl[3 * r] * Y[0] + pl[3 * r + 1] * Y[1] + pl[3 * r + 2]
                                 : pl[3 * r] * Y[2] + pl[3 * r + 1] * Y[3] + pl[3 * r + 2];
        if (res > 0) mx[1] = fmax(mx[1], res);
      }
    }
  }
  block_reduce<1, false>(s, c.red());
  block_reduce<2, true>(mx, c.red());
  const double err_kin = s[0] / (double)Nt;
  return err_kin < 1e-2 && mx[1] < 1e-1 && mx[0] < 1e-1;
}

template <int RC>
__device__ __noinline__ bool is_feasible_cold(CtxShared *s) {
  Ctx c = make_ctx(s, 0.0, 1.0);
  return is_feasible(c, s->P);
}
// initial corridors from the guess (float centres) or regeneration from the new iterate (double centres)
template <int RC>
__device__ __noinline__ int corridors_cold(CtxShared *s, bool from_solution, double *stage, int stage_doubles) {
  Ctx c = make_ctx(s, 0.0, 1.0);
  const int Nt = c.Nt(), NT = c.NT();
  if (from_solution) return agent_corridors(c, s->P, c.sol(), c.sol() + NT, c.sol() + 2 * NT, stage, stage_doubles, true, nullptr);
  return agent_corridors(c, s->P, c.guess(), c.guess() + Nt, c.guess() + 2 * Nt, stage, stage_doubles, false, nullptr);
}

// ===================================================================
// the kernel
// ===================================================================
// Register budget follows the block size (one thread per time step): horizons <= 128 run with up to
// 255 registers (2 CTAs/SM), <= 256 with 255 (1 CTA/SM), longer ones with 128.
template <int RC>
__device__ __forceinline__ void refine_body(const DevBatch &B, const DevOut &O, const csdo_params &P,
                                            const Layout &LY, double *scratch, int *queue, const QueueState &ST) {
  extern __shared__ double smem[];
  __shared__ int s_agent;
  __shared__ int s_flag;
  __shared__ CtxShared cs;
  Ctx c;
  c.s = &cs;
  const int NT = LY.NT;
  double *slot = scratch + (size_t)blockIdx.x * LY.slot_doubles;
  if (threadIdx.x == 0) {
    cs.NT = LY.NT; cs.KP = 4 * LY.KMAX;
    cs.P = P;
#if CSDO_TU == 1
    static_assert(one_warp(RC), "translation unit 1 holds the one-warp variants only");
#ifndef CSDO_OW_DIRECT
    cs.fn_solve = reinterpret_cast<void *>(*(volatile OwSolveFn *)&g_ow_solve[(LY.tier & 2) ? 0 : 1]);
    cs.fn_factor = reinterpret_cast<void *>(*(volatile OwFactorFn *)&g_ow_factor[(LY.tier & 2) ? 0 : 1]);
#endif
    cs.fn_sweep = nullptr;
#else
    static_assert(!one_warp(RC), "translation unit 2 holds the CTA-wide variants only");
    cs.fn_solve = reinterpret_cast<void *>(*(volatile BandSolveFn *)&g_band_solve[RC]);
    cs.fn_sweep = reinterpret_cast<void *>(*(volatile PbcrSweepFn *)&g_pbcr_sweep[(LY.tier & 2) ? 0 : 1]);
#endif
    cs.x = smem + LY.o_x; cs.xt = smem + LY.o_xt; cs.rhs = smem + LY.o_rhs; cs.D = smem + LY.o_D;
    cs.carry = smem + LY.o_carry; cs.red = smem + LY.o_red;
    const bool rows_glob = LY.tier & 1;
    cs.rows_glob = rows_glob;
    cs.ros = rows_glob ? slot + LY.g_ro : smem + LY.o_ro;
    cs.cfgs = cs.ros + RO_COUNT * NT;
    cs.Es = rows_glob ? slot + LY.g_E : smem + LY.o_E;
    cs.ws = (rows_glob && !LY.w_smem) ? slot + LY.g_w : smem + LY.o_w;
    cs.w_shared = !rows_glob || LY.w_smem;
    cs.pstart = reinterpret_cast<int *>(smem + LY.o_pstart);
    cs.pm.L = (LY.tier & 2) ? slot + LY.g_L : smem + LY.o_L;
    cs.l_shared = !(LY.tier & 2);
    cs.pm.S = smem + LY.o_sinv;
    cs.pm.g = cs.carry;                    // carry, xt are free while a solve runs
    cs.pm.y = cs.xt;
    cs.pm.xs = cs.xt + 6 * (NT / kPM);
    if (one_warp(RC)) {   // the same areas, laid out for the one-warp solver (make_layout sizes them)
      cs.bm.L6 = cs.pm.L;
      cs.bm.dinv = cs.bm.L6 + 36 * NT + kSkewPad;
      cs.bm.Sinv = smem + LY.o_sinv;
      cs.bm.sv = cs.bm.Sinv + kL2Doubles;
      cs.bm.tab = cs.skew_tab;
      cs.bm.G = cs.xt;  // xt, rhs and carry are contiguous (16 NT doubles) and free while a factorization runs
    }
    cs.cur = slot + LY.g_cur; cs.sol = slot + LY.g_sol; cs.dy = slot + LY.g_dy;
    cs.pl_glob = slot + LY.g_pl; cs.pl_smem = smem + LY.o_pl; cs.KS = LY.KS;
    cs.pc_glob = slot + LY.g_pc; cs.pc_smem = smem + LY.o_pc; cs.pc_cap = LY.PC;
  }

#ifdef CSDO_DEV_TIMERS
  for (int k = 0; k < 8; ++k) c.ph[k] = 0;
#endif
  __syncthreads();
  for (;;) {
    if (threadIdx.x == 0) {
      int a_next = -1;
      const int idx = atomicAdd(queue + kQHead, 1);
      if (idx < ST.cap) {
        volatile int *slot = ST.items + idx;
        volatile int *remaining = queue + kQRemaining;
        // Liveness: a CTA only waits while the queue is empty and some SQP loop is still running; the CTAs
        // that run those loops never wait for anybody, so they always reach their re-enqueue / retire step,
        // which either fills this slot or drops `remaining` to 0.  That holds whether or not the whole grid
        // is co-resident (a CTA that is not resident yet holds no slot).  The time bound below is a safety
        // net against a lost update only: two minutes of wall clock, far beyond any refine.
        unsigned long long t_start = 0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
        while ((a_next = *slot) < 0) {  // the slot is filled by the CTA that re-enqueues an agent
          if (*remaining <= 0) break;    // every SQP loop is finished: nothing will be appended any more
          __nanosleep(200);
          unsigned long long t_now;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_now));
          if (t_now - t_start > 120000000000ull) { atomicExch(queue + kQError, 1); break; }
        }
      }
      s_agent = a_next;
    }
    __syncthreads();
    const int a = s_agent;
    __syncthreads();
    if (a < 0) break;
    __threadfence();  // the agent's state was published by another CTA before its queue slot was written
    // instance of this agent: last i with inst_agent_ptr[i] <= a
    int lo = 0, hi = B.n_inst;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (B.inst_agent_ptr[mid] <= a) lo = mid; else hi = mid;
    }
    const int inst = lo;
    const int Nt = B.inst_nt[inst];
    const int64_t off = B.agent_off[a];
    if (threadIdx.x == 0) {
      cs.Nt = Nt;
      if (one_warp(RC)) {
        fill_skew_table(Nt, cs.skew_tab);
        cs.solver_warp = a % ((blockDim.x + 31) >> 5);
      }
      cs.K = B.plane_ptr[a + 1] - B.plane_ptr[a];
      if (cs.K > LY.KMAX) {  // the caller understated max_planes: stay inside the scratch slot and flag it (csdo_sync)
        atomicExch(queue + kQError, 2);
        cs.K = LY.KMAX;
      }
      cs.pl = cs.K <= cs.KS ? cs.pl_smem : cs.pl_glob;
      cs.plane_t = B.plane_t + B.plane_ptr[a];
      cs.plane_abc = B.plane_abc + (size_t)12 * B.plane_ptr[a];
      cs.No = B.obs_ptr[inst + 1] - B.obs_ptr[inst];
      cs.obs = B.obs + (size_t)3 * B.obs_ptr[inst];
      cs.dimx = B.inst_dims[2 * inst]; cs.dimy = B.inst_dims[2 * inst + 1];
      cs.guess = B.guess + 6 * off;
      cs.corr = O.corridors + 8 * off;
    }
    __syncthreads();
    double *traj = O.traj + 6 * off;
    // planes of each step: planes are sorted by t (inter_agent_cons.cc:26-31)
    for (int tt = c.tid(); tt <= Nt; tt += c.nthr()) {
      int l2 = 0, h2 = c.K();
      while (l2 < h2) { const int mid = (l2 + h2) >> 1; if (c.plane_t()[mid] < tt) l2 = mid + 1; else h2 = mid; }
      c.pstart()[tt] = l2;
    }
    // solution0 and the frozen trust centre (dsqp_solver.cc:56-63); cfg (utils.cc:115-120).  Later passes
    // continue from the iterate the previous pass left in the trajectory output.
    const bool first_pass = __ldcg(&O.sqp_iters[a]) == 0;  // (zeroed before the launch)
    if (c.active()) {
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        double v = first_pass ? c.guess()[k * Nt + c.t()] : __ldcg(&traj[k * Nt + c.t()]);
        if (k >= 4 && !c.has_next()) v = 0.0;
        c.cur()[k * NT + c.t()] = v;
        c.sol()[k * NT + c.t()] = v;
      }
      c.ros()[RO_TRX * NT + c.t()] = c.guess()[0 * Nt + c.t()];
      c.ros()[RO_TRY * NT + c.t()] = c.guess()[1 * Nt + c.t()];
    }
    if (threadIdx.x == 0) {
      c.cfgs()[0] = c.guess()[0]; c.cfgs()[1] = c.guess()[Nt - 1];
      c.cfgs()[2] = c.guess()[Nt]; c.cfgs()[3] = c.guess()[2 * Nt - 1];
      c.cfgs()[4] = c.guess()[2 * Nt]; c.cfgs()[5] = c.guess()[3 * Nt - 1];
    }
    if (threadIdx.x == 0) s_flag = 0;
    __syncthreads();
    PH_T0();
    if (first_pass) {
      // calcCorridors (dsqp_solver.cc:1154) on float disc centres
      if (corridors_cold<RC>(c.s, false, smem + LY.o_x, LY.o_carry - LY.o_x)) s_flag = 1;
      __syncthreads();
      if (threadIdx.x == 0 && s_flag) atomicAnd(&O.inst_static_legal[inst], 0);
      __syncthreads();
    }

    // `while (delta > th && iter_count < max_iter)` (dsqp_solver.cc:99-253): one iteration per visit
    // (see QueueState)
    const double th = P.delta_solution_threshold;
    int iter_count = first_pass ? 0 : __ldcg(&O.sqp_iters[a]);
    int status = first_pass ? 1 : __ldcg(&O.status[a]);
    int admm = first_pass ? 0 : __ldcg(&O.admm_iters[a]), nfac = first_pass ? 0 : __ldcg(&O.n_factor[a]);
    bool finished = true;
    __syncthreads();
    PH_ADD(0);
    while (iter_count < P.max_iter) {
      assemble_rows_cold<RC>(c.s);
      PH_ADD(1);
      const QpOut q = solve_qp<RC>(c, P, LY);
      PH_RESET();
      status = q.status; admm += q.iters; nfac += q.n_factor;
      double s[1] = {0.0};
      if (c.active()) {
        const int nv = nvar(c);
#pragma unroll
        for (int k = 0; k < 6; ++k)
          if (k < nv) { const double d = c.sol()[k * NT + c.t()] - c.cur()[k * NT + c.t()]; s[0] += d * d; }
      }
      block_reduce<1, false>(s, c.red());
      const double delta = s[0];
      iter_count++;
      if (iter_count > P.max_iter / 2 && is_feasible_cold<RC>(c.s)) { finished = true; break; }
      if (c.active())
#pragma unroll
        for (int k = 0; k < 6; ++k) c.cur()[k * NT + c.t()] = c.sol()[k * NT + c.t()];
      __syncthreads();
      if (!P.fixed_corridor) {
        PH_ADD(7);
        corridors_cold<RC>(c.s, true, smem + LY.o_x, LY.o_carry - LY.o_x);
        __syncthreads();
        PH_ADD(0);
      }
      finished = !(delta > th && iter_count < P.max_iter);
      if (finished || iter_count < kMaxVisits - 1) break;  // one SQP iteration per visit (see kMaxVisits)
    }
    // extractSingleSolutionVec2OptRes + per-agent records
    double ob[1] = {0.0};
    if (c.active()) {
#pragma unroll
      for (int k = 0; k < 6; ++k) traj[k * Nt + c.t()] = c.sol()[k * NT + c.t()];
      if (c.has_next()) {
        const double w0 = c.sol()[5 * NT + c.t()];
        ob[0] = 0.5 * w0 * w0;
        if (c.t() + 1 < Nt - 1) {
          const double dv = c.sol()[4 * NT + c.t() + 1] - c.sol()[4 * NT + c.t()];
          ob[0] += 0.5 * dv * dv;
        }
      }
    }
    block_reduce<1, false>(ob, c.red());
    if (threadIdx.x == 0) {
      O.status[a] = status; O.sqp_iters[a] = iter_count; O.n_qp[a] = iter_count;
      O.admm_iters[a] = admm; O.n_factor[a] = nfac; O.objective[a] = ob[0];
    }
    // publish the agent's state, then either retire it or append it to the queue for its next iteration
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      if (finished) {
        atomicSub(queue + kQRemaining, 1);
      } else {
        const int slot = atomicAdd(queue + kQTail, 1);
        if (slot < ST.cap) *(volatile int *)(ST.items + slot) = a;
        else atomicExch(queue + kQError, 1);
      }
    }
    __syncthreads();
    PH_ADD(7);
  }
#ifdef CSDO_DEV_TIMERS
  if (threadIdx.x == 0) {
    unsigned long long *prof = reinterpret_cast<unsigned long long *>(queue + 2);
    for (int k = 0; k < 8; ++k) atomicAdd(prof + k, (unsigned long long)c.ph[k]);
  }
#endif
}

template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
dsqp_refine_kernel(const DevBatch B, const DevOut O, const csdo_params P, const Layout LY, double *scratch,
                   int *queue, const QueueState ST) {
  // register class: 0 = 255 registers, 1 = 168 (three warps per sub-partition), 2 = 128 (four); +3 for the
  // block sizes that use the one-warp solver
  constexpr int kWarps = MAXT / 32 * MINB;
  constexpr int kRegs = (kWarps > 12) ? 2 : ((kWarps > 8) ? 1 : 0);
  static_assert(MAXT > kOneWarpMaxNT || kRegs < 2, "no 128-register variant of the one-warp solver");
  refine_body<(MAXT <= kOneWarpMaxNT) ? 3 + kRegs : kRegs>(B, O, P, LY, scratch, queue, ST);
}

using RefineKernel = void (*)(const DevBatch, const DevOut, const csdo_params, const Layout, double *, int *,
                              const QueueState);
// lean: the variant compiled for one more resident CTA per SM.  Registers are allocated per SM
// sub-partition (16384 each): 9 or 10 resident warps put 3 on one sub-partition, i.e. <= 168 per thread.
#if CSDO_TU == 1
RefineKernel pick_kernel_short(int block, bool lean) {   // block sizes <= 96
  if (block <= 64) return dsqp_refine_kernel<64, 4>;
  return lean ? dsqp_refine_kernel<96, 3> : dsqp_refine_kernel<96, 2>;
}
void read_debug_counters_short(unsigned long long *out32) {
  cudaMemcpyFromSymbol(out32, g_dbg, 32 * sizeof(unsigned long long));
  unsigned long long z[32] = {0};
  cudaMemcpyToSymbol(g_dbg, z, sizeof(z));
}
#else
RefineKernel pick_kernel_short(int block, bool lean);       // translation unit 1
void read_debug_counters_short(unsigned long long *out32);
static RefineKernel pick_kernel(int block, bool lean) {
  if (block <= kOneWarpMaxNT) return pick_kernel_short(block, lean);
  if (block <= 128) return dsqp_refine_kernel<128, 2>;
  if (block <= 160 && lean) return dsqp_refine_kernel<160, 2>;
  // (tried: <192,2> at 168 registers and <256,2> at 128 with the band factor in global scratch, i.e. two CTAs
  // per SM for long horizons -- 53.8 k vs 55.3 k QP/s at Nt = 190 and 37.5 k vs 48.3 k at Nt = 256: the factor
  // read from L2 inside the sweeps costs more than the second CTA brings)
  if (block <= 256) return dsqp_refine_kernel<256, 1>;
  return dsqp_refine_kernel<512, 1>;
}

// status aggregation of SolverDSQP (dsqp_solver.cc:1224-1243), one thread per instance
__global__ void aggregate_status_kernel(const DevBatch B, const DevOut O) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B.n_inst) return;
  bool any = false;
  int worst = 2;
  for (int a = B.inst_agent_ptr[i]; a < B.inst_agent_ptr[i + 1]; ++a) {
    const int s = O.status[a];
    if (abs(s) > 1) { any = true; if (abs(s) > worst) worst = s; }
  }
  O.inst_status[i] = any ? worst : 1;
}

__global__ void fill_int_kernel(int *p, int n, int v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// corridors only (csdo_corridors): one CTA per agent
__global__ void corridors_kernel(const DevBatch B, const csdo_params P, int double_centres, double *corridors,
                                 int *box_status, int *inst_static_legal) {
  __shared__ int s_flag;
  const int a = blockIdx.x;
  int lo = 0, hi = B.n_inst;
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (B.inst_agent_ptr[mid] <= a) lo = mid; else hi = mid; }
  const int inst = lo, Nt = B.inst_nt[inst];
  const int64_t off = B.agent_off[a];
  __shared__ CtxShared cs;
  Ctx c;
  c.s = &cs;
  if (threadIdx.x == 0) {
    cs.Nt = Nt;
    cs.No = B.obs_ptr[inst + 1] - B.obs_ptr[inst];
    cs.obs = B.obs + (size_t)3 * B.obs_ptr[inst];
    cs.dimx = B.inst_dims[2 * inst]; cs.dimy = B.inst_dims[2 * inst + 1];
    cs.corr = corridors + 8 * off;
  }
  const double *g = B.guess + 6 * off;
  if (threadIdx.x == 0) s_flag = 0;
  __syncthreads();
  if (agent_corridors(c, P, g, g + Nt, g + 2 * Nt, nullptr, 0, double_centres != 0,
                      box_status ? box_status + 4 * off : nullptr))
    s_flag = 1;
  __syncthreads();
  if (threadIdx.x == 0 && s_flag) atomicAnd(&inst_static_legal[inst], 0);
}

// ===================================================================
// host-side launchers (called from csdo_api.cpp through dsqp_launch.h)
// ===================================================================
Layout make_layout(int NT, int KMAX, int tier, int KS, int PC, bool w_smem) {
  // tier bit 0: per-step row data / row scaling / row state in global scratch instead of shared memory
  // tier bit 1: band factor in global scratch
  Layout l{};
  l.NT = NT; l.KMAX = KMAX; l.tier = tier; l.KS = KS; l.PC = PC; l.w_smem = (tier & 1) && w_smem;
  const bool rows_glob = tier & 1, band_glob = tier & 2;
  int o = 0;
  auto take = [&](int n) { int r = o; o += (n + 1) & ~1; return r; };
  l.o_x = take(6 * NT); l.o_D = take(6 * NT); l.o_xt = take(6 * NT); l.o_rhs = take(6 * NT);
  l.o_carry = take(4 * NT);
  if (!rows_glob) {
    l.o_ro = take(RO_COUNT * NT + 8);
    l.o_E = take(16 * NT);
    l.o_w = take(16 * NT);
  }
  l.o_red = take((((NT < 64 ? 64 : NT) + 31) / 32 + 1) * N_COUNT);  // one row per warp + the result row
  l.o_pstart = take((NT + 2 + 1) / 2);
  const bool ow = NT <= kOneWarpMaxNT;   // the kernels with block size <= 96 use the one-warp solver's storage
  const int L_doubles = ow ? kLw * 6 * NT + kSkewPad : pbcr_L_doubles(NT);
  l.o_sinv = take(ow ? kL2Doubles + 3 * kMaxNs : pbcr_S_doubles(NT));
  l.o_L = band_glob ? 0 : take(L_doubles);
  l.o_pl = take(PL_COUNT * 4 * KS);
  l.o_pc = take(PC);
  if (l.w_smem) l.o_w = take(16 * NT);  // tier 1 with room left: the rows' ADMM state w back on chip (read + written every pass)
  l.smem_doubles = o;
  size_t g = 0;
  auto gtake = [&](size_t n) { size_t r = g; g += (n + 1) & ~(size_t)1; return r; };
  l.g_cur = gtake(6 * (size_t)NT); l.g_sol = gtake(6 * (size_t)NT); l.g_dy = gtake(16 * (size_t)NT);
  l.g_pl = gtake((size_t)PL_COUNT * 4 * KMAX);
  l.g_pc = gtake((size_t)6 * KMAX);
  l.g_L = gtake((size_t)L_doubles);
  l.g_ro = gtake((size_t)RO_COUNT * NT + 8); l.g_E = gtake(16 * (size_t)NT); l.g_w = gtake(16 * (size_t)NT);
  l.slot_doubles = g;
  return l;
}

// queue and per-agent counters before the launch
__global__ void queue_init_kernel(int n, int n_agents, const int *order, int *items, int cap, int *ctrl,
                                  int *sqp_iters) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cap) items[i] = i < n ? (order ? order[i] : i) : -1;
  if (i < n) sqp_iters[order ? order[i] : i] = 0;  // (an agent's first visit is recognised by sqp_iters == 0)
  if (i == 0) { ctrl[kQHead] = 0; ctrl[kQTail] = n; ctrl[kQRemaining] = n; ctrl[kQError] = 0; }
}

size_t refine_queue_bytes(int n_agents, const csdo_params &P) {
  const size_t visits = (size_t)(P.max_iter > 1 ? (P.max_iter < kMaxVisits ? P.max_iter : kMaxVisits) : 1);
  return ((size_t)n_agents * visits + 64) * sizeof(int);
}

cudaError_t launch_init_outputs(const DevBatch &B, const DevOut &O, cudaStream_t stream) {
  const int fb = 256;
  fill_int_kernel<<<(B.n_inst + fb - 1) / fb, fb, 0, stream>>>(O.inst_static_legal, B.n_inst, 1);
  return cudaGetLastError();
}

cudaError_t launch_refine(const DevBatch &B, const DevOut &O, const csdo_params &P, const Layout &LY,
                          double *scratch, int *queue, void *queue_items, int grid, int block, bool lean,
                          cudaStream_t stream, int *n_launches, bool init_outputs, bool aggregate) {
  const int smem = LY.smem_doubles * 8;
  RefineKernel kern = pick_kernel(block, lean);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  if (const char *co = getenv("CSDO_CARVEOUT"))  // developer knob: shared-memory carve-out in percent
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(co));
  const int fb = 256, n = (B.n_active > 0 && B.agent_order) ? B.n_active : B.n_agents;
  QueueState qs{static_cast<int *>(queue_items), (int)(refine_queue_bytes(n, P) / sizeof(int))};
  e = cudaMemsetAsync(queue, 0, 2048, stream);
  if (e != cudaSuccess) return e;
  if (init_outputs) fill_int_kernel<<<(B.n_inst + fb - 1) / fb, fb, 0, stream>>>(O.inst_static_legal, B.n_inst, 1);
  queue_init_kernel<<<(qs.cap + fb - 1) / fb, fb, 0, stream>>>(n, B.n_agents, B.agent_order, qs.items, qs.cap, queue,
                                                               O.sqp_iters);
  kern<<<grid, block, smem, stream>>>(B, O, P, LY, scratch, queue, qs);
  if (aggregate) aggregate_status_kernel<<<(B.n_inst + fb - 1) / fb, fb, 0, stream>>>(B, O);
  if (n_launches) *n_launches = 2 + (init_outputs ? 1 : 0) + (aggregate ? 1 : 0);
  return cudaGetLastError();
}

cudaError_t launch_aggregate_status(const DevBatch &B, const DevOut &O, cudaStream_t stream) {
  const int fb = 256;
  aggregate_status_kernel<<<(B.n_inst + fb - 1) / fb, fb, 0, stream>>>(B, O);
  return cudaGetLastError();
}

cudaError_t launch_corridors(const DevBatch &B, const csdo_params &P, int double_centres, double *corridors,
                             int *box_status, int *inst_static_legal, cudaStream_t stream) {
  const int fb = 256;
  fill_int_kernel<<<(B.n_inst + fb - 1) / fb, fb, 0, stream>>>(inst_static_legal, B.n_inst, 1);
  corridors_kernel<<<B.n_agents, 128, 0, stream>>>(B, P, double_centres, corridors, box_status,
                                                   inst_static_legal);
  return cudaGetLastError();
}

int refine_occupancy(int block, int smem_bytes, bool lean) {
  RefineKernel kern = pick_kernel(block, lean);
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess) return 0;
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, block, smem_bytes) != cudaSuccess) return 0;
  return n;
}

void read_debug_counters(unsigned long long *out32) {   // both translation units' counters, summed
  cudaMemcpyFromSymbol(out32, g_dbg, 32 * sizeof(unsigned long long));
  unsigned long long z[32] = {0};
  cudaMemcpyToSymbol(g_dbg, z, sizeof(z));
  read_debug_counters_short(z);
  for (int i = 0; i < 32; ++i) out32[i] += z[i];
}

int refine_kernel_regs(int block, bool lean) {
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, pick_kernel(block, lean)) != cudaSuccess) return -1;
  return fa.numRegs;
}

#endif  // CSDO_TU == 2 (else branch of the pick_kernel block)

}  // namespace csdo
