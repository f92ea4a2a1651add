/*
 * csdo_dsqp.h -- C ABI of the B200-native DSQP refine stage.
 *
 * This is the drop-in boundary for ONE path of YangSVM/CSDOTrajectoryPlanning:
 * the decentralized-SQP refine stage that the reference runs inside the
 * constructor of `SolverDSQP` (reference sqp/dsqp_solver.h:24-47,
 * sqp/dsqp_solver.cc:1133-1248), plus the two pre-process functions that feed
 * it (reference sqp/inter_agent_cons.cc:12-140) and the corridor builder it
 * calls (reference sqp/corridor.cc:124-248).
 *
 * The reference has no C ABI (it is one C++ executable); these entry points
 * are what a cgo/ctypes/C++ binding for that path binds.  The C++ shim that
 * keeps the reference's own class interface lives in include/csdo/dsqp_solver.h.
 *
 * Conventions: plain pointers and sizes, caller owns every buffer, int return
 * codes (0 = ok), no exceptions cross the boundary, one CUDA stream per
 * handle, a handle is not thread-safe.  There is NO CPU fallback: every
 * compute entry point fails with CSDO_ERR_CUDA when no sm_100 device is
 * usable.
 *
 * Batch layout (all FP64 unless noted).  A batch is n_inst independent
 * instances; agents are numbered globally, instance i owns agents
 * [inst_agent_ptr[i], inst_agent_ptr[i+1]).  All agents of an instance share
 * the horizon inst_nt[i] (reference inter_agent_cons.cc:320-345 pads to the
 * longest path).  Agent a stores its horizon at time-step offset
 * agent_off[a] (agent_off[a+1]-agent_off[a] == Nt of its instance):
 *   guess / traj : 6 planes of Nt doubles at 6*agent_off[a]:
 *                  x[Nt] y[Nt] yaw[Nt] steer[Nt] v[Nt] w[Nt]
 *                  (v,w: Nt-1 valid, last entry ignored on input, 0 on output;
 *                  the field names follow OptimizeResult, sqp/common.h:14-22,
 *                  with w == d_steer)
 *   corridors    : 8 planes of Nt doubles at 8*agent_off[a], in the field
 *                  order of `Corridor` (sqp/corridor.h:8-11):
 *                  xf_min xf_max yf_min yf_max xr_min xr_max yr_min yr_max
 * Planes: CSR over agents, plane_ptr[n_agents+1]; plane k has time plane_t[k]
 * and 12 doubles in the member order of `InterPlane`
 * (sqp/inter_agent_cons.h:47-53): a,b,c of f2f, f2r, r2f, r2r.  Within an
 * agent planes are sorted by t (the reference's push order,
 * inter_agent_cons.cc:26-31,136-137).
 * Obstacles: CSR over instances, 3 doubles (x, y, r) each, in the iteration
 * order of the caller's container (reference: std::unordered_set<Location>).
 */
#ifndef CSDO_DSQP_H_
#define CSDO_DSQP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSDO_OK 0
#define CSDO_ERR_INVALID 1     /* bad argument / inconsistent batch */
#define CSDO_ERR_CUDA 2        /* no usable device or a CUDA call failed */
#define CSDO_ERR_UNSUPPORTED 3 /* horizon too long for the kernel's layout */
#define CSDO_ERR_NOMEM 4

/* OSQP status codes that can leave the QP step (osqp/constants.h of 0.6.x). */
#define CSDO_QP_SOLVED 1
#define CSDO_QP_SOLVED_INACCURATE 2
#define CSDO_QP_PRIMAL_INFEASIBLE_INACCURATE 3
#define CSDO_QP_MAX_ITER_REACHED (-2)
#define CSDO_QP_PRIMAL_INFEASIBLE (-3)
#define CSDO_QP_NON_CVX (-7)
#define CSDO_QP_UNSOLVED (-10)

/*
 * Parameters.  Vehicle constants follow `Constants` (common/motion_planning.h
 * :9-67, values from config.yaml via common/motion_planning.cc:54-109), the
 * SQP block follows `QpParm` (sqp/common.h:39-52, sqp/utils.cc:34-59), the OSQP
 * block holds what osqp_set_default_settings() of OSQP 0.6.x leaves in force
 * at sqp/dsqp_solver.cc:480-487.
 */
typedef struct csdo_params {
  /* vehicle */
  double f2x;       /* front disc offset,  (3 LF - LB)/4 */
  double r2x;       /* rear disc offset,   (LF - 3 LB)/4 */
  double rv;        /* disc radius */
  double WB;        /* wheel base */
  double steer_max; /* atan(WB / r), dsqp_solver.cc:1178 */
  double LF, LB, car_width; /* rectangle, used by the SAT legality flag only */
  /* QpParm */
  double r_trust;
  double max_omega;
  double max_v;
  double delta_solution_threshold;
  double dt;
  int32_t max_iter;       /* SQP iterations (config max_iter) */
  int32_t osqp_max_iter;  /* ADMM iterations per QP */
  int32_t fixed_corridor; /* bool */
  /* OSQP 0.6.x settings */
  int32_t adaptive_rho_interval; /* PINNED (reference: wall-clock dependent); 25 */
  int32_t scaling;               /* Ruiz passes, 10 */
  int32_t check_termination;     /* 25 */
  int32_t adaptive_rho;          /* 1 */
  double rho, sigma, alpha;
  double eps_abs, eps_rel, eps_prim_inf, eps_dual_inf;
  double adaptive_rho_tolerance;
  /* corridor growth (sqp/corridor.h:72 defaults) */
  double box_ds;    /* 0.1 */
  double box_limit; /* 10.0 */
} csdo_params;

typedef struct csdo_batch {
  int32_t n_inst;
  int32_t n_agents;
  const int32_t *inst_agent_ptr; /* [n_inst+1] */
  const int32_t *inst_nt;        /* [n_inst] */
  const double *inst_dims;       /* [n_inst][2] dimx, dimy */
  const int32_t *obs_ptr;        /* [n_inst+1] */
  const double *obs;             /* [sum No][3] */
  const int64_t *agent_off;      /* [n_agents+1], time-step offsets */
  const double *guess;           /* [6 * agent_off[n_agents]] */
  const int32_t *plane_ptr;      /* [n_agents+1] */
  const int32_t *plane_t;        /* [sum K] */
  const double *plane_abc;       /* [sum K][12] */
  const int32_t *agent_order;    /* [n_agents] optional initial order of the
                                    work queue (a permutation, longest first
                                    balances the GPU best); NULL = the library
                                    decides.  An agent is handed out for one SQP
                                    iteration at a time and re-enqueued until
                                    its loop ends. */
  int32_t n_active;              /* 0 = every agent.  > 0 (needs agent_order):
                                    only the first n_active agents of agent_order
                                    are refined / get planes; the others' results
                                    are left untouched and their plane lists are
                                    empty.  This is the agent-partitioned multi-GPU
                                    mode: every rank holds the whole batch (the
                                    plane build needs every agent's guess) and
                                    works on its own agents. */
} csdo_batch;

typedef struct csdo_result {
  double *traj;               /* [6 * total steps] */
  double *corridors;          /* [8 * total steps] */
  int32_t *status;            /* [n_agents] OSQP code of the agent's LAST QP */
  int32_t *sqp_iters;         /* [n_agents] == SolverDSQP::num_iterations */
  int32_t *n_qp;              /* [n_agents] QPs solved (== sqp_iters) */
  int32_t *admm_iters;        /* [n_agents] ADMM iterations summed over QPs */
  int32_t *n_factor;          /* [n_agents] KKT factorizations summed over QPs */
  double *objective;          /* [n_agents] objective of the last QP at its output */
  int32_t *inst_status;       /* [n_inst] == SolverDSQP::getSolverStatus() */
  int32_t *inst_static_legal; /* [n_inst] == get_initial_static_legal() */
} csdo_result;

typedef struct csdo_handle csdo_handle;

/* Fills the values the reference computes from its shipped config.yaml. */
void csdo_default_params(csdo_params *p);

/* Library / build identification ("csdo-dsqp-b200 <ver> sm_100a"). */
const char *csdo_version(void);

/* Create a solver bound to CUDA device `device`.  Replaces nothing in the
 * reference (it has no device); owns the stream, scratch and work queue. */
int csdo_create(const csdo_params *params, int device, csdo_handle **out);
void csdo_destroy(csdo_handle *h);
const char *csdo_last_error(const csdo_handle *h);

/*
 * Refine a batch: for every instance what `SolverDSQP::SolverDSQP` does
 * (dsqp_solver.cc:1133-1248): initial corridors, per-agent SQP loop with an
 * OSQP-style ADMM per QP, corridor regeneration, feasibility early exit,
 * status aggregation.  HOST pointers; copies in, runs, copies out, blocks.
 */
int csdo_refine(csdo_handle *h, const csdo_batch *in, csdo_result *out);

/* Same, but every pointer in both structs is a DEVICE pointer (inputs already
 * resident in HBM, results stay there).  max_nt / max_planes: the largest
 * horizon and the largest per-agent plane count of the batch (host ints, so
 * no device->host read is needed to size the launch).  Enqueued on
 * `cuda_stream` (a cudaStream_t; NULL = the handle's stream); returns without
 * synchronizing. */
int csdo_refine_device(csdo_handle *h, const csdo_batch *in, csdo_result *out,
                       int max_nt, int max_planes, void *cuda_stream);

/* csdo_refine_device with HOST copies of the small per-instance arrays, so that
 * the library can group the agents by horizon without reading device memory:
 * one launch per horizon class (block size 64, 96, 128, ... steps), longest
 * first, all on `cuda_stream` -- in a batch of mixed horizons the short agents
 * then run in their own launch shape (3-4 CTAs per SM) instead of the longest
 * agent's.  csdo_refine (host
 * buffers) always works this way.  host_inst_nt [n_inst], host_inst_agent_ptr
 * [n_inst+1]: host copies of in->inst_nt / in->inst_agent_ptr; host_order
 * [n_order]: the agents to refine in processing order (NULL / 0: all agents,
 * longest horizon first); in->agent_order / in->n_active are ignored.  Returns
 * without synchronizing.  Results are those of csdo_refine_device up to the
 * rounding of the solver variant a horizon class selects. */
int csdo_refine_device_hinted(csdo_handle *h, const csdo_batch *in, csdo_result *out,
                              int max_planes, const int32_t *host_inst_nt,
                              const int32_t *host_inst_agent_ptr,
                              const int32_t *host_order, int32_t n_order,
                              void *cuda_stream);

/* The launch plan csdo_refine / csdo_refine_device_hinted use for a set of
 * horizons (host only, no device needed): agents grouped by horizon class,
 * longest class first; classes with fewer than min_count agents merged into the
 * next larger class of the same solver family (<= 96 steps / above).  Writes
 * the grouped agent ids to order_out[n_agents] and per bucket its longest
 * horizon and agent count; returns the number of buckets (<= 8) or -1. */
int csdo_plan_horizon_buckets(int32_t n_agents, const int32_t *agent_nt, int32_t min_count,
                              int32_t *order_out, int32_t *bucket_nt, int32_t *bucket_count,
                              int32_t max_buckets);

/* SolverDSQP's status aggregation (dsqp_solver.cc:1224-1243) over ALL agents of
 * every instance from out->status into out->inst_status (DEVICE pointers).
 * csdo_refine* run it themselves; it is exported for the agent-partitioned
 * mode, where the statuses of the other ranks' agents arrive by all-gather. */
int csdo_aggregate_status_device(csdo_handle *h, const csdo_batch *in,
                                 csdo_result *out, void *cuda_stream);

/* Waits for the last csdo_refine_device of this handle and reports what only
 * the device knows: CSDO_ERR_INVALID if an agent had more planes than the
 * max_planes passed (its result is then invalid), CSDO_ERR_CUDA if the work
 * queue stalled.  A handle owns ONE scratch area and work queue: calls on the
 * same handle must be ordered on one stream (or separated by csdo_sync); use
 * one handle per concurrent stream. */
int csdo_sync(csdo_handle *h);

/* Kernel-launch and timing facts of the last refine on this handle
 * (launches: kernels enqueued; smem_bytes/block/grid: the DSQP kernel's
 * configuration; tier: 0 all-shared, 1 read-only rows in global, 2 band
 * factor in global). */
typedef struct csdo_launch_info {
  int32_t launches;
  int32_t grid, block, smem_bytes, tier, ctas_per_sm;
} csdo_launch_info;
int csdo_last_launch(const csdo_handle *h, csdo_launch_info *info);

/*
 * Initial corridors only (reference calcCorridors, sqp/corridor.cc:164-248,
 * or, with double_centres != 0, the regeneration of
 * SolverDSQP::updateCorridor, dsqp_solver.cc:818-872).  Uses in->guess x,y,yaw.
 * HOST pointers.  corridors: [8*steps]; box_status: [2*steps][2] int32
 * (success, initial_status) front then rear per step, may be NULL.
 */
int csdo_corridors(csdo_handle *h, const csdo_batch *in, int double_centres,
                   double *corridors, int32_t *box_status,
                   int32_t *inst_static_legal);

/*
 * Pre-process: neighbour pairs + separating planes
 * (findNeighborPairsByTrustRegion, inter_agent_cons.cc:12-49 and
 * calcEqualInterPlanes, :71-140).  Two calls: count fills plane_ptr
 * [n_agents+1] and inst_inter_legal[n_inst] (the function's return value);
 * the caller then allocates plane_t / plane_abc for plane_ptr[n_agents]
 * planes and calls fill.  HOST pointers; in->plane_* are ignored.
 */
int csdo_planes_count(csdo_handle *h, const csdo_batch *in, int32_t *plane_ptr,
                      int32_t *inst_inter_legal);
int csdo_planes_fill(csdo_handle *h, const csdo_batch *in,
                     const int32_t *plane_ptr, int32_t *plane_t,
                     double *plane_abc);
/* Same, and plane_partner[k] (may be NULL) receives the global id of the other
 * agent of plane k: the pair list of findNeighborPairsByTrustRegion
 * (inter_agent_cons.cc:12-49) is {(plane_t[k], a, partner[k]) : a < partner[k]}
 * sorted by (t, a, partner). */
int csdo_planes_fill_partners(csdo_handle *h, const csdo_batch *in,
                              const int32_t *plane_ptr, int32_t *plane_t,
                              double *plane_abc, int32_t *plane_partner);

/* calcEqualInterPlanes (inter_agent_cons.cc:71-140) for an explicit pair list:
 * pairs[p] = {t, i, j} with global agent ids i, j of one instance.  Every pair
 * pushes one plane to agent i and one to agent j, in list order.  HOST pointers;
 * plane_ptr [n_agents+1] out, plane_t [2 n_pairs], plane_abc [2 n_pairs][12]. */
int csdo_planes_from_pairs(csdo_handle *h, const csdo_batch *in, int64_t n_pairs,
                           const int32_t *pairs, int32_t *plane_ptr,
                           int32_t *plane_t, double *plane_abc);

/* Device-resident pre-process (every pointer a DEVICE pointer, enqueued on
 * cuda_stream, NULL = the handle's stream; in->plane_* ignored).
 * count: step_off [total_steps+1] receives the index of the first plane of
 * every (agent, step) (exclusive scan on the device; the last entry is the
 * total), plane_ptr [n_agents+1], inst_inter_legal [n_inst].  If total_planes
 * (a HOST pointer) is non-NULL the call synchronizes and stores the plane
 * count there so that the caller can size plane_t / plane_abc.
 * fill: plane_t [total], plane_abc [total][12], plane_partner [total] or NULL. */
int csdo_planes_count_device(csdo_handle *h, const csdo_batch *in, int64_t total_steps,
                             int32_t *step_off, int32_t *plane_ptr,
                             int32_t *inst_inter_legal, int64_t *total_planes,
                             void *cuda_stream);
int csdo_planes_fill_device(csdo_handle *h, const csdo_batch *in,
                            const int32_t *step_off, int32_t *plane_t,
                            double *plane_abc, int32_t *plane_partner,
                            void *cuda_stream);

/*
 * Measurement helper (no reference counterpart): sustained FP64 FMA throughput
 * of `device` in TFLOP/s from a register-resident DFMA microbenchmark.  It is
 * the roofline denominator of the DSQP kernel, which is FP64-pipe bound, not
 * HBM or tensor-core bound.
 */
int csdo_measure_fp64_peak(int device, double *tflops_out);

#ifdef __cplusplus
}
#endif
#endif /* CSDO_DSQP_H_ */
