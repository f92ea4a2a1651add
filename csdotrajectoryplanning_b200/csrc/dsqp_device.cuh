// Device-side building blocks of the DSQP refine kernel (sm_100a).
//
// One CTA refines one agent: the whole per-agent SQP loop of the reference's
// SolverDSQP::calcIndividualSQP (sqp/dsqp_solver.cc:36-269) runs inside the
// CTA, one thread per time step of the horizon.  The agent QP is never
// materialised as a sparse matrix: every constraint row of time step t touches
// only the 6 unknowns of step t (x,y,yaw,steer,v,w) and, for the 4 kinematic
// rows, x,y,yaw,steer of step t+1, so thread t regenerates its rows from a few
// per-step coefficients (visit_rows below).  In time-major order the reduced
// KKT matrix P + sigma I + A' diag(rho) A is symmetric positive definite with
// half-bandwidth 6; its LDL' factor lives in shared memory.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "csdo_dsqp.h"

namespace csdo {

constexpr double kRhoMin = 1e-6, kRhoMax = 1e6, kRhoEqOverIneq = 1e3, kRhoTol = 1e-4;
constexpr double kMinScaling = 1e-4, kMaxScaling = 1e4, kOsqpInfty = 1e30;
constexpr int kBand = 6;        // half bandwidth of the reduced KKT in time-major order
constexpr int kLw = kBand + 1;  // stored entries per row: 1/d_i and 6 sub-diagonals

// read-only per-step planes (stride NT)
enum Ro : int {
  RO_SN = 0, RO_CS, RO_A1, RO_A2, RO_A3, RO_B3, RO_KR0, RO_KR1, RO_KR2,
  RO_CL0, RO_CL1, RO_CL2, RO_CL3, RO_CU0, RO_CU1, RO_CU2, RO_CU3, RO_TRX, RO_TRY, RO_COUNT
};
// per plane-row arrays (stride KP = 4*KMAX)
enum Pl : int { PL_A = 0, PL_B, PL_G, PL_U, PL_E, PL_W, PL_COUNT };
// local unknown indices inside visit_rows: own step 0..5, next step x,y,yaw,steer 6..9
enum Var : int { VX = 0, VY, VP, VS, VV, VW, NX, NY, NP, NS };

struct BandMem {
  double *L6, *dinv;  // [(6t+k)*6 + d-1], [6t+k]; generic pointers (shared or global)
  double *Sinv;       // dense inverse of the separator Schur complement (shared)
  double *sv;         // 3 * kMaxNs scratch: g, x_sep, pivot row (shared)
  double *G;          // factor-time scratch: per partition GCC[21] GBB[21] GBC[36] (shared)
};

struct Ctx {
  int Nt, NT, K, KP, No, tid, nthr, t;  // t = this thread's time step (tid), valid if tid < Nt
  bool active, has_next;
  // shared-memory vectors, SoA with stride NT: v[k*NT + t]
  double *x, *xt, *rhs, *D, *carry, *w, *E, *red;
  double *cfgw, *cfgE;  // 6 start/goal rows (contiguous after w / E)
  int *pstart;          // [Nt+1] first plane of each step
  BandMem bm;           // band factor storage (band_solver.cuh)
  double *ro;           // RO_COUNT planes, generic pointer
  // per-CTA global scratch
  double *cur, *sol, *dy, *pl;
  // batch views of this agent
  const double *guess;      // 6 planes, stride Nt
  const double *plane_abc;  // [K][12]
  const int *plane_t;       // [K]
  const double *obs;        // [No][3]
  double *corr;             // 8 planes, stride Nt (output array doubles as the live corridor)
  double dimx, dimy;
  double cfg[6];
  double rho, c;  // current rho and Ruiz cost scaling
};

__device__ __forceinline__ double limit_scaling(double v) {
  v = v < kMinScaling ? 1.0 : v;
  return v > kMaxScaling ? kMaxScaling : v;
}
__device__ __forceinline__ double clipd(double v, double lo, double hi) {
  // c_min(c_max(v, l), u) of OSQP's project()
  double a = (v > lo) ? v : lo;
  return (a < hi) ? a : hi;
}

// ---- block reductions (deterministic): N running values per thread ----
template <int N, bool IS_MAX>
__device__ __forceinline__ void block_reduce(double (&v)[N], double *red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double a = v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double b = __shfl_xor_sync(0xffffffffu, a, o);
      a = IS_MAX ? fmax(a, b) : a + b;
    }
    if (lane == 0) red[warp * N + i] = a;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double a = red[i];
    for (int wv = 1; wv < nw; ++wv) a = IS_MAX ? fmax(a, red[wv * N + i]) : a + red[wv * N + i];
    v[i] = a;
  }
  __syncthreads();
}

// ---- the rows of time step t (reference sqp/dsqp_solver.cc:646-1129) ----
// f.row<NC>(i0,c0,i1,c1,i2,c2,i3,c3, l, u, w, E): NC coefficients on local
// unknowns i*, raw bounds l,u, the row's ADMM state w and Ruiz factor E.
template <class F>
__device__ __forceinline__ void visit_rows(const Ctx &c, const csdo_params &P, F &f) {
  const int NT = c.NT, t = c.t;
  const double sn = c.ro[RO_SN * NT + t], cs = c.ro[RO_CS * NT + t];
  if (c.has_next) {
    // calcKineConstraint :646-744, lb = ub = -C
    const double kr0 = c.ro[RO_KR0 * NT + t], kr1 = c.ro[RO_KR1 * NT + t], kr2 = c.ro[RO_KR2 * NT + t];
    f.template row<4>(VX, 1.0, VP, c.ro[RO_A1 * NT + t], VV, P.dt * cs, NX, -1.0, kr0, kr0, c.w[0 * NT + t], c.E[0 * NT + t]);
    f.template row<4>(VY, 1.0, VP, c.ro[RO_A2 * NT + t], VV, P.dt * sn, NY, -1.0, kr1, kr1, c.w[1 * NT + t], c.E[1 * NT + t]);
    f.template row<4>(VP, 1.0, VS, c.ro[RO_A3 * NT + t], VV, c.ro[RO_B3 * NT + t], NP, -1.0, kr2, kr2, c.w[2 * NT + t], c.E[2 * NT + t]);
    f.template row<3>(VS, 1.0, VW, P.dt * 1.0, NS, -1.0, 0, 0.0, -0.0, -0.0, c.w[3 * NT + t], c.E[3 * NT + t]);
  }
  // calcCfgConstraint :746-788 (cfg = x0,xN,y0,yN,yaw0,yawN)
  if (t == 0) {
    f.template row<1>(VX, 1.0, 0, 0.0, 0, 0.0, 0, 0.0, c.cfg[0], c.cfg[0], c.cfgw[0], c.cfgE[0]);
    f.template row<1>(VY, 1.0, 0, 0.0, 0, 0.0, 0, 0.0, c.cfg[2], c.cfg[2], c.cfgw[2], c.cfgE[2]);
    f.template row<1>(VP, 1.0, 0, 0.0, 0, 0.0, 0, 0.0, c.cfg[4], c.cfg[4], c.cfgw[4], c.cfgE[4]);
  }
  if (t == c.Nt - 1) {
    f.template row<1>(VX, 1.0, 0, 0.0, 0, 0.0, 0, 0.0, c.cfg[1], c.cfg[1], c.cfgw[1], c.cfgE[1]);
    f.template row<1>(VY, 1.0, 0, 0.0, 0, 0.0, 0, 0.0, c.cfg[3], c.cfg[3], c.cfgw[3], c.cfgE[3]);
    f.template row<1>(VP, 1.0, 0, 0.0, 0, 0.0, 0, 0.0, c.cfg[5], c.cfg[5], c.cfgw[5], c.cfgE[5]);
  }
  // calcCorridorConstraint :874-968: D = [I,0,-f2x sin; 0,I,f2x cos; I,0,-r2x sin; 0,I,r2x cos]
  f.template row<2>(VX, 1.0, VP, -P.f2x * sn, 0, 0.0, 0, 0.0, c.ro[RO_CL0 * NT + t], c.ro[RO_CU0 * NT + t], c.w[4 * NT + t], c.E[4 * NT + t]);
  f.template row<2>(VY, 1.0, VP, P.f2x * cs, 0, 0.0, 0, 0.0, c.ro[RO_CL1 * NT + t], c.ro[RO_CU1 * NT + t], c.w[5 * NT + t], c.E[5 * NT + t]);
  f.template row<2>(VX, 1.0, VP, -P.r2x * sn, 0, 0.0, 0, 0.0, c.ro[RO_CL2 * NT + t], c.ro[RO_CU2 * NT + t], c.w[6 * NT + t], c.E[6 * NT + t]);
  f.template row<2>(VY, 1.0, VP, P.r2x * cs, 0, 0.0, 0, 0.0, c.ro[RO_CL3 * NT + t], c.ro[RO_CU3 * NT + t], c.w[7 * NT + t], c.E[7 * NT + t]);
  // calcTrustRegionConstraint :970-994 (centre = initial guess, all SQP iterations)
  {
    const double trx = c.ro[RO_TRX * NT + t], try_ = c.ro[RO_TRY * NT + t];
    f.template row<1>(VX, 1.0, 0, 0.0, 0, 0.0, 0, 0.0, -P.r_trust + trx, P.r_trust + trx, c.w[8 * NT + t], c.E[8 * NT + t]);
    f.template row<1>(VY, 1.0, 0, 0.0, 0, 0.0, 0, 0.0, -P.r_trust + try_, P.r_trust + try_, c.w[9 * NT + t], c.E[9 * NT + t]);
  }
  // calcMaxCtrlAndSteerConstraint :996-1039
  if (c.has_next) {
    f.template row<1>(VV, 1.0, 0, 0.0, 0, 0.0, 0, 0.0, -P.max_v, P.max_v, c.w[10 * NT + t], c.E[10 * NT + t]);
    f.template row<1>(VW, 1.0, 0, 0.0, 0, 0.0, 0, 0.0, -P.max_omega, P.max_omega, c.w[11 * NT + t], c.E[11 * NT + t]);
  }
  f.template row<1>(VS, 1.0, 0, 0.0, 0, 0.0, 0, 0.0, -P.steer_max, P.steer_max, c.w[12 * NT + t], c.E[12 * NT + t]);
  // calcInterVehicleConstraint :1097-1129: 4 rows per plane of this step, l = -inf
  const int k0 = c.pstart[t], k1 = c.pstart[t + 1];
  for (int r = 4 * k0; r < 4 * k1; ++r) {
    f.template row<3>(VX, c.pl[PL_A * c.KP + r], VY, c.pl[PL_B * c.KP + r], VP, c.pl[PL_G * c.KP + r], 0, 0.0,
                      -INFINITY, c.pl[PL_U * c.KP + r], c.pl[PL_W * c.KP + r], c.pl[PL_E * c.KP + r]);
  }
}

// rho of a row from its scaled bounds (OSQP set_rho_vec; "loose" rows cannot occur)
__device__ __forceinline__ double row_rho(double ls, double us, double rho) {
  return (us - ls < kRhoTol) ? kRhoEqOverIneq * rho : rho;
}

}  // namespace csdo
