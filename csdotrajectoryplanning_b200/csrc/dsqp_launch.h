// Host <-> kernel interface of the DSQP refine path (internal to the library).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "csdo_dsqp.h"

namespace csdo {

constexpr int kMaxThreads = 512;  // one thread per time step => horizon <= 512

struct DevBatch {  // device pointers, same meaning as csdo_batch
  int n_inst, n_agents;
  const int *inst_agent_ptr, *inst_nt;
  const double *inst_dims;
  const int *obs_ptr;
  const double *obs;
  const int64_t *agent_off;
  const double *guess;
  const int *plane_ptr, *plane_t;
  const double *plane_abc;
  const int *agent_order;  // may be null
  int n_active;            // 0: all agents; else the first n_active entries of agent_order
};

struct DevOut {  // device pointers, same meaning as csdo_result
  double *traj, *corridors;
  int *status, *sqp_iters, *n_qp, *admm_iters, *n_factor;
  double *objective;
  int *inst_status, *inst_static_legal;
};

// Work queue of the refine kernel (one persistent launch).  An agent is handed out for ONE SQP
// iteration at a time and, unless its SQP loop is finished, appended to the queue again (round-robin
// time slicing): the agents whose QPs run into the ADMM iteration limit (a few percent of the agents,
// ~40 % of the work, up to 10 QPs x 400 iterations) then advance together with everything else instead
// of occupying single CTAs at the end of the launch.  The state between two visits of an agent is its
// trajectory / corridor / counter output (FP64 / int), read back around the (non-coherent) L1.
//   ctrl[kQHead] next queue slot to hand out, ctrl[kQTail] next free slot, ctrl[kQRemaining] agents
//   whose SQP loop is not finished, ctrl[kQError] set if a waiting CTA gave up (never expected)
constexpr int kQHead = 0, kQTail = 20, kQRemaining = 21, kQError = 22;
// An agent is handed out at most kMaxVisits times: from its last visit on it runs to the end of its SQP
// loop (bounds the queue for large QpParm::max_iter; the default max_iter = 10 never gets there).
constexpr int kMaxVisits = 16;
struct QueueState {
  int *items;  // [cap] agent ids, -1 = not written yet
  int cap;     // n_agents x min(max(1, max_iter), kMaxVisits) + slack
};

// Shared-memory / scratch placement for one launch (offsets in doubles).
struct Layout {
  int NT, KMAX, tier, smem_doubles;
  int KS;  // planes per agent that fit the shared-memory plane area
  int PC;  // doubles of the shared-memory plane-contribution buffer (visit_planes)
  bool w_smem;  // tier 1 only: the fixed rows' state w kept in shared memory
  int o_x, o_xt, o_rhs, o_D, o_carry, o_red, o_pstart, o_L, o_sinv, o_pl, o_pc, o_ro, o_E, o_w;
  size_t g_cur, g_sol, g_dy, g_pl, g_pc, g_L, g_ro, g_E, g_w, slot_doubles;
};

Layout make_layout(int NT, int KMAX, int tier, int KS, int PC, bool w_smem);
int refine_occupancy(int block, int smem_bytes, bool lean);
int refine_kernel_regs(int block, bool lean);
void read_debug_counters(unsigned long long *out16);

// queue_items: refine_queue_bytes() of device memory
size_t refine_queue_bytes(int n_agents, const csdo_params &P);
// init_outputs: set inst_static_legal to 1 first; aggregate: run the status aggregation afterwards (a bucketed
// refine does both once, around its per-bucket launches)
cudaError_t launch_refine(const DevBatch &B, const DevOut &O, const csdo_params &P, const Layout &LY,
                          double *scratch, int *queue, void *queue_items, int grid, int block, bool lean,
                          cudaStream_t stream, int *n_launches, bool init_outputs = true, bool aggregate = true);
cudaError_t launch_init_outputs(const DevBatch &B, const DevOut &O, cudaStream_t stream);
cudaError_t launch_aggregate_status(const DevBatch &B, const DevOut &O, cudaStream_t stream);
cudaError_t launch_corridors(const DevBatch &B, const csdo_params &P, int double_centres, double *corridors,
                             int *box_status, int *inst_static_legal, cudaStream_t stream);

// neighbour pairs + planes (planes_kernel.cu)
cudaError_t launch_planes_count(const DevBatch &B, const csdo_params &P, int64_t total_steps, int *step_cnt,
                                int *inst_inter_legal, cudaStream_t stream);
cudaError_t launch_planes_fill(const DevBatch &B, const csdo_params &P, const int *step_off, int *plane_t,
                               double *plane_abc, int *plane_partner, cudaStream_t stream);
cudaError_t launch_planes_from_pairs(const DevBatch &B, const csdo_params &P, int64_t n_pairs, const int *pairs,
                                     const int *pos, int *plane_t, double *plane_abc, cudaStream_t stream);
// exclusive scan of the per-step plane counts on the device (step_cnt [steps + 1], in place)
cudaError_t launch_plane_offsets(const DevBatch &B, int64_t steps, int *step_cnt, int *tile_sum, int *plane_ptr,
                                 cudaStream_t stream);

}  // namespace csdo
