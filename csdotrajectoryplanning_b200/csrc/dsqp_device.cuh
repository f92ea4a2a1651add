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
// half-bandwidth 6; its factor lives in shared memory (pbcr_solver.cuh).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "csdo_dsqp.h"

namespace csdo {
// developer counters: cycle timers, compiled in with -DCSDO_DEV_TIMERS (make DEV=-DCSDO_DEV_TIMERS) and
// printed by csdo_refine when CSDO_PROFILE is set.  Thread 0 of every CTA accumulates clock64 deltas.
static __device__ unsigned long long g_dbg[32];   // (one copy per translation unit)
}
#ifdef CSDO_DEV_TIMERS
#define DBG_INIT() long long dbg_t_ = clock64()
#define DBG_ACC(i) do { if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&csdo::g_dbg[i], (unsigned long long)(t_ - dbg_t_)); dbg_t_ = t_; } } while (0)
#else
#define DBG_INIT() do {} while (0)
#define DBG_ACC(i) do {} while (0)
#endif

#include "pbcr_solver.cuh"

namespace csdo {

constexpr double kRhoMin = 1e-6, kRhoMax = 1e6, kRhoEqOverIneq = 1e3, kRhoTol = 1e-4;
constexpr double kMinScaling = 1e-4, kMaxScaling = 1e4, kOsqpInfty = 1e30;
constexpr int kBand = 6;        // half bandwidth of the reduced KKT in time-major order
constexpr int kLw = kBand + 1;  // one-warp solver: stored entries per row (1/d_i and 6 sub-diagonals)

// geometry of the one-warp solver used for horizons <= 96 steps (band_solver.cuh)
constexpr int kMaxP = 16;                 // horizon partitions (lanes of the solver warp)
constexpr int kMaxNs = 6 * (kMaxP - 1);   // separator unknowns
constexpr int kL2Tinv = 4 * 18 * 18, kL2R = 18 * 18, kL2B = 6 * 36;  // level-2 inverses and couplings
constexpr int kL2Doubles = kL2Tinv + kL2R + kL2B;
constexpr int kSkewPad = 256;  // extra doubles at the end of the L6 area (<= 16 partitions x 14 doubles)
constexpr int kOneWarpMaxNT = 96;         // block sizes up to this use the one-warp solver

struct BandMem {
  double *L6, *dinv;  // [(6t+k)*6 + d-1], [6t+k]; generic pointers (shared or global)
  double *Sinv;       // level-2 inverses of the separator system, kL2Doubles (shared)
  double *sv;         // 3 * kMaxNs scratch: g, x_sep, z / pivot rows (shared)
  double *G;          // factor-time scratch: per partition GCC[21] GBB[21] GBC[36] = 78 * kMaxP doubles (shared)
  const int *tab;     // bank-skew table of the partitions (shared context)
};

// read-only per-step planes (stride NT)
enum Ro : int {
  RO_SN = 0, RO_CS, RO_A1, RO_A2, RO_A3, RO_B3, RO_KR0, RO_KR1, RO_KR2,
  RO_CL0, RO_CL1, RO_CL2, RO_CL3, RO_CU0, RO_CU1, RO_CU2, RO_CU3, RO_TRX, RO_TRY, RO_COUNT
};
// per plane-row record (AoS, 6 doubles = 3 LDS.128): pl[row * 6 + field]
enum Pl : int { PL_A = 0, PL_B, PL_G, PL_U, PL_E, PL_W, PL_COUNT };
// local unknown indices inside visit_rows: own step 0..5, next step x,y,yaw,steer 6..9
enum Var : int { VX = 0, VY, VP, VS, VV, VW, NX, NY, NP, NS };


// CTA-uniform context of the agent being refined.  It lives in SHARED memory (written by thread 0
// between barriers): keeping two dozen pointers per thread in registers starved the row passes.
struct CtxShared {
  int Nt, NT, K, KP, No, KS;
  bool l_shared, rows_glob, w_shared;
  // shared-memory vectors, SoA with stride NT: v[k*NT + t]
  double *x, *xt, *rhs, *D, *carry, *red;
  int *pstart;   // [Nt+1] first plane of each step
  double *ros;   // read-only per-step row data, RO_COUNT planes
  double *cfgs;  // 6 start/goal pins
  double *Es;    // Ruiz row scaling of the fixed rows, 16 planes (read-only during the ADMM loop)
  double *ws;    // ADMM row state w = z_hat + y/rho of the fixed rows, 16 planes
  PbcrMem pm;    // reduced-KKT factor storage (pbcr_solver.cuh), horizons > 96
  csdo_params P; // copy of the parameters for the out-of-line (cold) phases
  void *fn_solve; // band solve entry point (indirect call, see dsqp_kernel.cu)
  void *fn_sweep; // the sweeps alone (developer variants)
  // per-CTA global scratch
  double *cur, *sol, *dy;
  double *pl;    // plane rows of this agent: shared memory when they fit (K <= KS), else global scratch
  double *pl_smem, *pl_glob;
  double *pc_smem, *pc_glob;  // per-plane contributions of a plane-major pass (visit_planes)
  int pc_cap;                 // doubles of pc_smem
  // batch views of this agent
  const double *guess;      // 6 planes, stride Nt
  const double *plane_abc;  // [K][12]
  const int *plane_t;       // [K]
  const double *obs;        // [No][3]
  double *corr;             // 8 planes, stride Nt (output array doubles as the live corridor)
  double dimx, dimy;
  // one-warp solver (band_solver.cuh), horizons <= 96 (kept at the end: the offsets above are what the
  // register allocation of the row passes was tuned with)
  BandMem bm;
  int skew_tab[kMaxP];
  int solver_warp;
  void *fn_factor; // one-warp factorization entry point
};

// Per-thread handle: the shared context plus the few values that change inside a QP.
struct Ctx {
  CtxShared *s;
  double rho, c;    // current rho and Ruiz cost scaling (every thread holds the same value)
#ifdef CSDO_DEV_TIMERS
  long long ph[8];  // phase cycle counters (developer profiling)
#endif
#define CSDO_GET(type, name) __device__ __forceinline__ type name() const { return s->name; }
  CSDO_GET(int, Nt) CSDO_GET(int, NT) CSDO_GET(int, K) CSDO_GET(int, KP) CSDO_GET(int, No) CSDO_GET(int, KS)
  CSDO_GET(bool, l_shared) CSDO_GET(bool, rows_glob) CSDO_GET(bool, w_shared)
  CSDO_GET(double *, x) CSDO_GET(double *, xt) CSDO_GET(double *, rhs) CSDO_GET(double *, D)
  CSDO_GET(double *, carry) CSDO_GET(double *, red) CSDO_GET(int *, pstart) CSDO_GET(double *, ros)
  CSDO_GET(double *, cfgs) CSDO_GET(double *, Es) CSDO_GET(double *, ws)
  CSDO_GET(double *, cur) CSDO_GET(double *, sol) CSDO_GET(double *, dy) CSDO_GET(double *, pl)
  CSDO_GET(double *, pl_smem) CSDO_GET(double *, pl_glob) CSDO_GET(double *, pc_smem) CSDO_GET(double *, pc_glob)
  CSDO_GET(int, pc_cap)
  CSDO_GET(const double *, guess) CSDO_GET(const double *, plane_abc) CSDO_GET(const int *, plane_t)
  CSDO_GET(const double *, obs) CSDO_GET(double *, corr) CSDO_GET(double, dimx) CSDO_GET(double, dimy)
#undef CSDO_GET
  __device__ __forceinline__ const PbcrMem &pm() const { return s->pm; }
  __device__ __forceinline__ const BandMem &bm() const { return s->bm; }
  // one thread per time step
  __device__ __forceinline__ int tid() const { return threadIdx.x; }
  __device__ __forceinline__ int nthr() const { return blockDim.x; }
  __device__ __forceinline__ int t() const { return threadIdx.x; }
  __device__ __forceinline__ bool active() const { return (int)threadIdx.x < s->Nt; }
  __device__ __forceinline__ bool has_next() const { return (int)threadIdx.x < s->Nt - 1; }
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
// red: ((blockDim/32) + 1) * N doubles of shared memory.  Warp partials by shuffles, then thread i < N
// folds the partials of value i in warp order (compact code: this runs rarely and is fetch-bound).
template <int N, bool IS_MAX>
__device__ __forceinline__ void block_reduce(double (&v)[N], double *red) {
  __builtin_assume(__isShared(red));
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
  if ((int)threadIdx.x < N) {
    double a = red[threadIdx.x];
#pragma unroll 1
    for (int wv = 1; wv < nw; ++wv) a = IS_MAX ? fmax(a, red[wv * N + threadIdx.x]) : a + red[wv * N + threadIdx.x];
    red[nw * N + threadIdx.x] = a;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = red[nw * N + i];
  __syncthreads();
}

// ---- the rows of time step t (reference sqp/dsqp_solver.cc:646-1129) ----
// f.row<NC, EQ>(rid, i0,c0,i1,c1,i2,c2,i3,c3, l, u, w, E): row id (0..15 fixed rows, -1 plane rows),
// NC coefficients on local unknowns i*, raw bounds l,u (EQ: l == u by construction), the row's ADMM
// state w and Ruiz factor E.
// A functor declares which of w / E it modifies (kWriteW / kWriteE).  The per-step row data, E and w
// of the 13 every-step rows are read into registers up front and written back once at the end: with
// the values behind shared-memory references the compiler had to keep every load behind the previous
// row's store (possible aliasing) and the pass ran one row at a time.
struct RowRegs {
  double ro[RO_COUNT], E[13], w[13];
};

__device__ __forceinline__ void rows_load(const Ctx &c, RowRegs &R) {
  const int t = c.t(), NTs = c.NT();
  const double *ro_ = c.ros() + t, *Ep = c.Es() + t, *wp = c.ws() + t;
  if (c.rows_glob()) {  // global scratch (plain loads: the read-only part is reused from L1 across passes)
#pragma unroll
    for (int i = 0; i < RO_COUNT; ++i) R.ro[i] = ro_[i * NTs];
#pragma unroll
    for (int i = 0; i < 13; ++i) R.E[i] = Ep[i * NTs];
    if (c.w_shared()) {   // the state w alone is back in shared memory
      __builtin_assume(__isShared(wp));
#pragma unroll
      for (int i = 0; i < 13; ++i) R.w[i] = wp[i * NTs];
    } else {
#pragma unroll
      for (int i = 0; i < 13; ++i) R.w[i] = wp[i * NTs];
    }
  } else {
    __builtin_assume(__isShared(ro_)); __builtin_assume(__isShared(Ep)); __builtin_assume(__isShared(wp));
#pragma unroll
    for (int i = 0; i < RO_COUNT; ++i) R.ro[i] = ro_[i * NTs];
#pragma unroll
    for (int i = 0; i < 13; ++i) { R.E[i] = Ep[i * NTs]; R.w[i] = wp[i * NTs]; }
  }
}

template <bool WRITE_W, bool WRITE_E>
__device__ __forceinline__ void rows_store(const Ctx &c, const RowRegs &R) {
  const int t = c.t(), NTs = c.NT();
  double *Ep = c.Es() + t, *wp = c.ws() + t;
  if (c.rows_glob()) {
    if (WRITE_W && c.w_shared()) {
      __builtin_assume(__isShared(wp));
#pragma unroll
      for (int i = 0; i < 13; ++i) wp[i * NTs] = R.w[i];
    } else if (WRITE_W) {
#pragma unroll
      for (int i = 0; i < 13; ++i) wp[i * NTs] = R.w[i];
    }
#pragma unroll
    for (int i = 0; i < 13; ++i)
      if (WRITE_E) Ep[i * NTs] = R.E[i];
  } else {
    __builtin_assume(__isShared(Ep)); __builtin_assume(__isShared(wp));
#pragma unroll
    for (int i = 0; i < 13; ++i) {
      if (WRITE_W) wp[i * NTs] = R.w[i];
      if (WRITE_E) Ep[i * NTs] = R.E[i];
    }
  }
}

// the 13 every-step rows (+ 3 start/goal rows) of this thread's step, from registers
template <class F>
__device__ __forceinline__ void rows_apply(Ctx &c, const csdo_params &P, F &f, RowRegs &R) {
  const int t = c.t(), NTs = c.NT();
  double (&w)[13] = R.w;
  double (&E)[13] = R.E;
#define RO(i) R.ro[i]
  const double sn = RO(RO_SN), cs = RO(RO_CS);
  if (c.has_next()) {
    // calcKineConstraint :646-744, lb = ub = -C
    f.template row<4, true>(0, VX, 1.0, VP, RO(RO_A1), VV, P.dt * cs, NX, -1.0, RO(RO_KR0), RO(RO_KR0), w[0], E[0]);
    f.template row<4, true>(1, VY, 1.0, VP, RO(RO_A2), VV, P.dt * sn, NY, -1.0, RO(RO_KR1), RO(RO_KR1), w[1], E[1]);
    f.template row<4, true>(2, VP, 1.0, VS, RO(RO_A3), VV, RO(RO_B3), NP, -1.0, RO(RO_KR2), RO(RO_KR2), w[2], E[2]);
    f.template row<3, true>(3, VS, 1.0, VW, P.dt * 1.0, NS, -1.0, 0, 0.0, -0.0, -0.0, w[3], E[3]);
  }
  // calcCfgConstraint :746-788 (cfg = x0,xN,y0,yN,yaw0,yawN); rows 13..15 of the first/last step
  if (t == 0 || t == c.Nt() - 1) {
    const int e = (t == 0) ? 0 : 1;
    double *wc = c.ws() + 13 * NTs + t, *Ec = c.Es() + 13 * NTs + t;
    f.template row<1, true>(13, VX, 1.0, 0, 0.0, 0, 0.0, 0, 0.0, c.cfgs()[0 + e], c.cfgs()[0 + e], wc[0], Ec[0]);
    f.template row<1, true>(14, VY, 1.0, 0, 0.0, 0, 0.0, 0, 0.0, c.cfgs()[2 + e], c.cfgs()[2 + e], wc[NTs], Ec[NTs]);
    f.template row<1, true>(15, VP, 1.0, 0, 0.0, 0, 0.0, 0, 0.0, c.cfgs()[4 + e], c.cfgs()[4 + e], wc[2 * NTs], Ec[2 * NTs]);
  }
  // calcCorridorConstraint :874-968: D = [I,0,-f2x sin; 0,I,f2x cos; I,0,-r2x sin; 0,I,r2x cos]
  f.template row<2>(4, VX, 1.0, VP, -P.f2x * sn, 0, 0.0, 0, 0.0, RO(RO_CL0), RO(RO_CU0), w[4], E[4]);
  f.template row<2>(5, VY, 1.0, VP, P.f2x * cs, 0, 0.0, 0, 0.0, RO(RO_CL1), RO(RO_CU1), w[5], E[5]);
  f.template row<2>(6, VX, 1.0, VP, -P.r2x * sn, 0, 0.0, 0, 0.0, RO(RO_CL2), RO(RO_CU2), w[6], E[6]);
  f.template row<2>(7, VY, 1.0, VP, P.r2x * cs, 0, 0.0, 0, 0.0, RO(RO_CL3), RO(RO_CU3), w[7], E[7]);
  // calcTrustRegionConstraint :970-994 (centre = initial guess, all SQP iterations)
  f.template row<1>(8, VX, 1.0, 0, 0.0, 0, 0.0, 0, 0.0, -P.r_trust + RO(RO_TRX), P.r_trust + RO(RO_TRX), w[8], E[8]);
  f.template row<1>(9, VY, 1.0, 0, 0.0, 0, 0.0, 0, 0.0, -P.r_trust + RO(RO_TRY), P.r_trust + RO(RO_TRY), w[9], E[9]);
  // calcMaxCtrlAndSteerConstraint :996-1039
  if (c.has_next()) {
    f.template row<1>(10, VV, 1.0, 0, 0.0, 0, 0.0, 0, 0.0, -P.max_v, P.max_v, w[10], E[10]);
    f.template row<1>(11, VW, 1.0, 0, 0.0, 0, 0.0, 0, 0.0, -P.max_omega, P.max_omega, w[11], E[11]);
  }
  f.template row<1>(12, VS, 1.0, 0, 0.0, 0, 0.0, 0, 0.0, -P.steer_max, P.steer_max, w[12], E[12]);
#undef RO
}

template <class F>
__device__ __forceinline__ void visit_rows(Ctx &c, const csdo_params &P, F &f) {
  RowRegs R;
  rows_load(c, R);
  rows_apply(c, P, f, R);
  rows_store<F::kWriteW, F::kWriteE>(c, R);
}

// ---- the inter-vehicle rows (calcInterVehicleConstraint, dsqp_solver.cc:1097-1129), PLANE-MAJOR ----
// 4 rows per plane, l = -inf.  Thread i handles planes i, i + nthr, ... of the agent (an agent has about
// one plane per step on average but up to ~8 on a crowded step: handled by the step's own thread, the
// passes waited for the slowest thread and paid one L2 round trip per plane, serially).  Every plane is
// processed by a fresh copy g of the caller's functor: plane_begin loads the three unknowns the rows touch
// (x, y, yaw of the plane's step), the 4 rows run from registers, plane_emit writes what the rows
// contribute to the step (3 or 6 doubles) to the contribution buffer, and after a barrier the step's thread
// folds its planes' contributions in plane order (plane_absorb): deterministic, no atomics.
// IN selects the per-step input: the current iterate D x (xt), D x~ right after a solve (D, rhs), D itself.
enum PlaneIn : int { IN_NONE = 0, IN_XT, IN_DXT, IN_D };

template <int IN, class F>
__device__ __forceinline__ void visit_planes(Ctx &c, const csdo_params &P, F &f) {
  const int K = c.K();
  if (K == 0) return;  // CTA-uniform
  const int NT = c.NT();
  constexpr int NOUT = F::kPlaneOut;
  // contribution buffer: the solve scratch xt (6 NT doubles, idle in every pass that does not read the iterate
  // from it), else the dedicated shared-memory buffer, else global scratch
  double *pc = (IN != IN_XT && NOUT * K <= 6 * NT) ? c.xt() : ((NOUT * K <= c.pc_cap()) ? c.pc_smem() : c.pc_glob());
  const bool on_chip = c.pl() == c.pl_smem();
  for (int k = c.tid(); k < K; k += c.nthr()) {
    double2 *q2 = reinterpret_cast<double2 *>(c.pl() + (size_t)PL_COUNT * 4 * k);
    double v[4][PL_COUNT];
    if (on_chip) {
      __builtin_assume(__isShared(q2));
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int h = 0; h < PL_COUNT / 2; ++h) {
          const double2 d2 = q2[r * (PL_COUNT / 2) + h];
          v[r][2 * h] = d2.x;
          v[r][2 * h + 1] = d2.y;
        }
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int h = 0; h < PL_COUNT / 2; ++h) {
          const double2 d2 = q2[r * (PL_COUNT / 2) + h];
          v[r][2 * h] = d2.x;
          v[r][2 * h + 1] = d2.y;
        }
    }
    F g = f;
    double in3[3] = {0.0, 0.0, 0.0};
    if (IN != IN_NONE) {
      const int t = c.plane_t()[k];
      if (IN == IN_XT) {
        in3[0] = c.xt()[VX * NT + t]; in3[1] = c.xt()[VY * NT + t]; in3[2] = c.xt()[VP * NT + t];
      } else if (IN == IN_DXT) {
        in3[0] = c.D()[VX * NT + t] * c.rhs()[VX * NT + t];
        in3[1] = c.D()[VY * NT + t] * c.rhs()[VY * NT + t];
        in3[2] = c.D()[VP * NT + t] * c.rhs()[VP * NT + t];
      } else {
        in3[0] = c.D()[VX * NT + t]; in3[1] = c.D()[VY * NT + t]; in3[2] = c.D()[VP * NT + t];
      }
    }
    g.plane_begin(in3);
#pragma unroll
    for (int r = 0; r < 4; ++r)
      g.template row<3>(-1, VX, v[r][PL_A], VY, v[r][PL_B], VP, v[r][PL_G], 0, 0.0, -INFINITY, v[r][PL_U],
                        v[r][PL_W], v[r][PL_E]);
    if (F::kWriteW || F::kWriteE) {
#pragma unroll
      for (int r = 0; r < 4; ++r) q2[r * (PL_COUNT / 2) + PL_E / 2] = make_double2(v[r][PL_E], v[r][PL_W]);
    }
    if (NOUT > 0) g.plane_emit(pc + (size_t)NOUT * k);
    f.plane_merge(g);
  }
  if (NOUT > 0) {
    __syncthreads();
    if (c.active()) {
      const int k0 = c.pstart()[c.t()], k1 = c.pstart()[c.t() + 1];
      for (int k = k0; k < k1; ++k) f.plane_absorb(pc + (size_t)NOUT * k);
    }
  }
}

// rho of a row from its scaled bounds (OSQP set_rho_vec; "loose" rows cannot occur)
__device__ __forceinline__ double row_rho(double ls, double us, double rho) {
  return (us - ls < kRhoTol) ? kRhoEqOverIneq * rho : rho;
}

}  // namespace csdo
