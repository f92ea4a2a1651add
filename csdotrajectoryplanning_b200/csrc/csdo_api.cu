// C ABI of the DSQP refine path (include/csdo_dsqp.h): handle, device memory,
// host<->device staging, launch configuration.  No CPU fallback: without a
// usable sm_100 device every compute entry point returns CSDO_ERR_CUDA.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "csdo_dsqp.h"
#include "dsqp_launch.h"

using namespace csdo;

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
};

struct csdo_handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t last_stream = nullptr;  // stream of the last csdo_refine_device (csdo_sync)
  csdo_params P{};
  std::string err;
  int num_sms = 0, smem_limit = 0, smem_limit_sm = 0;
  DevBuf scratch, queue, step_cnt, pass_buf, tile_sum;
  // horizon buckets of one refine (run_refine_bucketed): the grouped processing order; `queue` holds one
  // control block of 2048 B per bucket (the buckets run one after the other and share everything else)
  DevBuf order_buf;
  int buckets_used = 1;
  std::vector<DevBuf> stage;  // staging buffers of the host-pointer entry points
  csdo_launch_info last{};
};

namespace {

// the entry points switch to the handle's device and restore the caller's current device on return
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev); else prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

constexpr int PL_BYTES_PER_PLANE = 6 * 4 * 8;
constexpr int kMaxBuckets = 8;
constexpr int kOneWarpMaxNTHost = 96;   // = kOneWarpMaxNT of dsqp_device.cuh (block sizes that use the one-warp solver)
constexpr int kQueueCtrlBytes = 2048;   // one work-queue control block

bool set_err(csdo_handle *h, const char *what, cudaError_t e) {
  if (e == cudaSuccess) return false;
  if (h) h->err = std::string(what) + ": " + cudaGetErrorString(e);
  return true;
}

int ensure(csdo_handle *h, DevBuf &b, size_t bytes) {
  if (bytes <= b.cap) return CSDO_OK;
  if (b.p) cudaFree(b.p);
  b.p = nullptr; b.cap = 0;
  size_t want = bytes + bytes / 8 + 256;
  if (set_err(h, "cudaMalloc", cudaMalloc(&b.p, want))) return CSDO_ERR_NOMEM;
  b.cap = want;
  return CSDO_OK;
}

struct Meta {
  int max_nt = 0, max_k = 0;
  int64_t steps = 0;
  int n_obs = 0, n_planes = 0;
};

int validate_host(csdo_handle *h, const csdo_batch *in, bool need_planes, Meta &m) {
  if (!in || in->n_inst < 0 || in->n_agents < 0) { h->err = "null or negative-sized batch"; return CSDO_ERR_INVALID; }
  if (in->n_agents == 0) return CSDO_OK;
  if (!in->inst_agent_ptr || !in->inst_nt || !in->inst_dims || !in->obs_ptr || !in->agent_off || !in->guess ||
      (need_planes && !in->plane_ptr)) { h->err = "null array in batch"; return CSDO_ERR_INVALID; }
  if (in->inst_agent_ptr[0] != 0 || in->inst_agent_ptr[in->n_inst] != in->n_agents) {
    h->err = "inst_agent_ptr does not cover the agents"; return CSDO_ERR_INVALID;
  }
  for (int i = 0; i < in->n_inst; ++i) {
    const int nt = in->inst_nt[i];
    if (nt < 3) { h->err = "horizon < 3"; return CSDO_ERR_INVALID; }
    if (nt > kMaxThreads) { h->err = "horizon exceeds 512 steps"; return CSDO_ERR_UNSUPPORTED; }
    m.max_nt = std::max(m.max_nt, nt);
    for (int a = in->inst_agent_ptr[i]; a < in->inst_agent_ptr[i + 1]; ++a)
      if (in->agent_off[a + 1] - in->agent_off[a] != nt) { h->err = "agent_off inconsistent with inst_nt"; return CSDO_ERR_INVALID; }
  }
  m.steps = in->agent_off[in->n_agents];
  if (in->obs_ptr[0] != 0) { h->err = "obs_ptr[0] != 0"; return CSDO_ERR_INVALID; }
  for (int i = 0; i < in->n_inst; ++i)
    if (in->obs_ptr[i + 1] < in->obs_ptr[i]) { h->err = "obs_ptr not monotonic"; return CSDO_ERR_INVALID; }
  m.n_obs = in->obs_ptr[in->n_inst];
  if (m.n_obs > 0 && !in->obs) { h->err = "null obstacle array"; return CSDO_ERR_INVALID; }
  if (need_planes) {
    if (in->plane_ptr[0] != 0) { h->err = "plane_ptr[0] != 0"; return CSDO_ERR_INVALID; }
    for (int a = 0; a < in->n_agents; ++a) {
      if (in->plane_ptr[a + 1] < in->plane_ptr[a]) { h->err = "plane_ptr not monotonic"; return CSDO_ERR_INVALID; }
      m.max_k = std::max(m.max_k, in->plane_ptr[a + 1] - in->plane_ptr[a]);
    }
    m.n_planes = in->plane_ptr[in->n_agents];
    if (m.n_planes > 0 && (!in->plane_t || !in->plane_abc)) { h->err = "null plane arrays"; return CSDO_ERR_INVALID; }
    // the kernel finds the planes of a step by binary search: sorted by t within an agent, 0 <= t < Nt
    for (int a = 0; a < in->n_agents; ++a) {
      const int nt = (int)(in->agent_off[a + 1] - in->agent_off[a]);
      int prev = 0;
      for (int k = in->plane_ptr[a]; k < in->plane_ptr[a + 1]; ++k) {
        const int t = in->plane_t[k];
        if (t < prev || t >= nt) { h->err = "plane_t must be sorted within an agent and lie in [0, Nt)"; return CSDO_ERR_INVALID; }
        prev = t;
      }
    }
  }
  if (in->n_active < 0 || in->n_active > in->n_agents || (in->n_active > 0 && !in->agent_order)) {
    h->err = "n_active needs agent_order and 0 <= n_active <= n_agents"; return CSDO_ERR_INVALID;
  }
  if (in->agent_order) {  // distinct ids in range: a duplicate would let two CTAs refine one agent
    std::vector<char> seen((size_t)in->n_agents, 0);
    for (int a = 0; a < (in->n_active > 0 ? in->n_active : in->n_agents); ++a) {
      const int v = in->agent_order[a];
      if (v < 0 || v >= in->n_agents || seen[v]) { h->err = "agent_order is not a permutation"; return CSDO_ERR_INVALID; }
      seen[v] = 1;
    }
  }
  return CSDO_OK;
}

// upload one host array into staging slot `slot`
template <class T>
int upload(csdo_handle *h, int slot, const T *src, size_t n, const T **dst) {
  if ((int)h->stage.size() <= slot) h->stage.resize(slot + 1);
  int rc = ensure(h, h->stage[slot], std::max<size_t>(n * sizeof(T), 16));
  if (rc) return rc;
  if (n && set_err(h, "cudaMemcpyAsync H2D",
                   cudaMemcpyAsync(h->stage[slot].p, src, n * sizeof(T), cudaMemcpyHostToDevice, h->stream)))
    return CSDO_ERR_CUDA;
  *dst = static_cast<const T *>(h->stage[slot].p);
  return CSDO_OK;
}
template <class T>
int devalloc(csdo_handle *h, int slot, size_t n, T **dst) {
  if ((int)h->stage.size() <= slot) h->stage.resize(slot + 1);
  int rc = ensure(h, h->stage[slot], std::max<size_t>(n * sizeof(T), 16));
  if (rc) return rc;
  *dst = static_cast<T *>(h->stage[slot].p);
  return CSDO_OK;
}
template <class T>
int download(csdo_handle *h, T *dst, const T *src, size_t n) {
  if (!dst || !n) return CSDO_OK;
  if (set_err(h, "cudaMemcpyAsync D2H", cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost, h->stream)))
    return CSDO_ERR_CUDA;
  return CSDO_OK;
}

int upload_batch(csdo_handle *h, const csdo_batch *in, const Meta &m, bool planes, DevBatch &B) {
  int rc;
  B.n_inst = in->n_inst; B.n_agents = in->n_agents;
  if ((rc = upload(h, 0, in->inst_agent_ptr, (size_t)in->n_inst + 1, &B.inst_agent_ptr))) return rc;
  if ((rc = upload(h, 1, in->inst_nt, (size_t)in->n_inst, &B.inst_nt))) return rc;
  if ((rc = upload(h, 2, in->inst_dims, (size_t)2 * in->n_inst, &B.inst_dims))) return rc;
  if ((rc = upload(h, 3, in->obs_ptr, (size_t)in->n_inst + 1, &B.obs_ptr))) return rc;
  if ((rc = upload(h, 4, in->obs, (size_t)3 * m.n_obs, &B.obs))) return rc;
  if ((rc = upload(h, 5, in->agent_off, (size_t)in->n_agents + 1, &B.agent_off))) return rc;
  if ((rc = upload(h, 6, in->guess, (size_t)6 * m.steps, &B.guess))) return rc;
  B.plane_ptr = nullptr; B.plane_t = nullptr; B.plane_abc = nullptr; B.agent_order = nullptr; B.n_active = 0;
  if (planes) {
    if ((rc = upload(h, 7, in->plane_ptr, (size_t)in->n_agents + 1, &B.plane_ptr))) return rc;
    if ((rc = upload(h, 8, in->plane_t, (size_t)m.n_planes, &B.plane_t))) return rc;
    if ((rc = upload(h, 9, in->plane_abc, (size_t)12 * m.n_planes, &B.plane_abc))) return rc;
  }
  return CSDO_OK;
}

DevBatch as_dev(const csdo_batch *in) {
  DevBatch B;
  B.n_inst = in->n_inst; B.n_agents = in->n_agents;
  B.inst_agent_ptr = in->inst_agent_ptr; B.inst_nt = in->inst_nt; B.inst_dims = in->inst_dims;
  B.obs_ptr = in->obs_ptr; B.obs = in->obs; B.agent_off = in->agent_off; B.guess = in->guess;
  B.plane_ptr = in->plane_ptr; B.plane_t = in->plane_t; B.plane_abc = in->plane_abc;
  B.agent_order = in->agent_order;
  B.n_active = in->agent_order ? in->n_active : 0;
  return B;
}

// stride / footprint class of a horizon: multiples of 32 up to the one-warp solver's limit, of 16 above
// Multiples of 32.  (Developer knob CSDO_NT_GRAN=16: strides in steps of 16 above 96 steps, so that horizons
// 129..144 fit two CTAs per SM (108 KB at NT = 144) -- measured slower on the real map set, 90.8 k vs 94.5 k QP/s:
// the extra classes are small and end up merged into longer ones.)
int launch_nt(int nt) {
  static const int gran = getenv("CSDO_NT_GRAN") ? std::max(16, atoi(getenv("CSDO_NT_GRAN")) & ~15) : 32;
  return nt <= kOneWarpMaxNTHost ? std::max(64, (nt + 31) & ~31) : ((nt + gran - 1) / gran * gran);
}

// configure + enqueue the refine kernels on device-resident data
// bucket: which queue control block the launch uses (csdo_sync reads every block's error flag);
// init_outputs / aggregate: see launch_refine
int run_refine(csdo_handle *h, const DevBatch &B, const DevOut &O, int max_nt, int max_k, cudaStream_t stream,
               int bucket = 0, bool init_outputs = true, bool aggregate = true, bool record = true) {
  if (B.n_agents == 0) return CSDO_OK;
  if (max_nt < 3) { h->err = "horizon < 3"; return CSDO_ERR_INVALID; }
  if (max_nt > kMaxThreads) { h->err = "horizon exceeds 512 steps"; return CSDO_ERR_UNSUPPORTED; }
  // NT = stride of the per-step arrays (what the shared-memory footprint scales with), block = NT rounded to warps
  const int NT = launch_nt(max_nt);
  const int KMAX = std::max(4, (max_k + 3) & ~3);
  // placement (tier bit 0: row data / scaling / state in global scratch, bit 1: band factor in global
  // scratch): everything in shared memory when that already allows the most resident CTAs, else the
  // row arrays move out (they are read once per pass into registers); the factor moves last
  const int block = std::max(64, (NT + 31) & ~31);
  Layout LY{};
  int occ = 0;
  bool lean = false;  // kernel variant compiled with fewer registers for one more resident CTA
  const char *force = getenv("CSDO_TIER");  // developer knob
  for (int tier = 0; tier <= 3; ++tier) {
    if (force && tier != atoi(force)) continue;
    if (!force && tier >= 2 && occ > 0) break;
    const Layout l = make_layout(NT, KMAX, tier, 0, 0, false);
    if (l.smem_doubles * 8 + 1024 > h->smem_limit) continue;
    const char *no_lean = getenv("CSDO_NO_LEAN");  // developer knob: never use the register-lean kernel variants
    for (int ln = 0; ln < ((no_lean && atoi(no_lean)) ? 1 : 2); ++ln) {
      const int o = refine_occupancy(block, l.smem_doubles * 8, ln);
      if (o > occ) { occ = o; LY = l; lean = ln; }
    }
  }
  if (occ >= 1) {
    // spend the shared memory that the chosen residency leaves free on (1) the per-plane contribution
    // buffer of the plane-major passes (3 doubles per plane in the ADMM loop), (2) the agents' plane rows
    // (192 B per plane): agents with K <= KS keep them on chip, the rest use global scratch
    // (a block may use sharedMemPerBlockOptin at most, which is 1 KB less than the SM's capacity)
    int budget = (std::min(h->smem_limit_sm / occ - 1024, h->smem_limit - 600) - LY.smem_doubles * 8) / 8;  // doubles
    budget = std::max(0, budget - 16);
    // Measured (B200, Nt = 256, 1 CTA/SM): spending the ~37 KB that are left on w or on plane records is no
    // faster than leaving them to L1 (the row data then hits L1 instead) -- 44.3 k vs 45.3 k QP/s -- so the
    // extras are only used on request
    if (!getenv("CSDO_EXTRA_SMEM")) budget = 0;
    // (0) tier 1: the fixed rows' state w (16 NT doubles, read and written in every pass) back on chip
    const bool w_smem = (LY.tier & 1) && budget >= 16 * NT && !getenv("CSDO_NO_W_SMEM");
    if (w_smem) budget -= 16 * NT;
    // (1) contribution buffer for agents whose planes do not fit the solve scratch (3 K > 6 NT)
    const int PC = (3 * KMAX > 6 * NT) ? (std::min(3 * KMAX, budget) & ~1) : 0;
    int KS = std::min(KMAX, std::max(0, (budget - PC) / (PL_BYTES_PER_PLANE / 8)));
    KS &= ~1;
    Layout l = make_layout(NT, KMAX, LY.tier, KS, PC, w_smem);
    if (refine_occupancy(block, l.smem_doubles * 8, lean) >= occ) LY = l;
    else if (getenv("CSDO_PROFILE")) fprintf(stderr, "[csdo] extra shared-memory layout rejected (%d B)\n", l.smem_doubles * 8);
  }
  if (occ < 1) { h->err = "horizon does not fit the kernel's shared-memory layout"; return CSDO_ERR_UNSUPPORTED; }
  if (const char *cap = getenv("CSDO_MAX_CTAS_PER_SM")) occ = std::max(1, std::min(occ, atoi(cap)));  // developer knob
  const int n_work = (B.n_active > 0 && B.agent_order) ? B.n_active : B.n_agents;
  const int grid = std::min(n_work, h->num_sms * occ);
  int rc;
  DevBuf &scratch = h->scratch;
  DevBuf &items = h->pass_buf;
  if ((rc = ensure(h, scratch, (size_t)grid * LY.slot_doubles * sizeof(double)))) return rc;
  if ((rc = ensure(h, h->queue, (size_t)kMaxBuckets * kQueueCtrlBytes))) return rc;
  if ((rc = ensure(h, items, refine_queue_bytes(n_work, h->P)))) return rc;
  int launches = 0;
  int *ctrl = reinterpret_cast<int *>(static_cast<char *>(h->queue.p) + (size_t)bucket * kQueueCtrlBytes);
  if (set_err(h, "launch_refine",
              launch_refine(B, O, h->P, LY, static_cast<double *>(scratch.p), ctrl, items.p, grid, block, lean, stream,
                            &launches, init_outputs, aggregate)))
    return CSDO_ERR_CUDA;
  if (record) {
    h->last.launches = launches;
    h->last.grid = grid; h->last.block = block; h->last.smem_bytes = LY.smem_doubles * 8;
    h->last.tier = LY.tier; h->last.ctas_per_sm = occ;
    h->buckets_used = 1;
  } else {
    h->last.launches += launches;
  }
  return CSDO_OK;
}

// ---- horizon buckets -------------------------------------------------------------------------------------
// One launch sizes its CTAs (threads, shared memory, solver variant) for the LONGEST horizon of what it is
// given: in a batch with mixed horizons (a real map set: 40 ... 250 steps) every short agent would run in the
// long agents' launch shape -- 1 CTA/SM instead of 3-4.  The agents are therefore grouped by horizon class
// (block size: 64, 96, 128, ... in steps of 32) and each class is refined by its own launch, longest horizons
// first, one after the other on the caller's stream (forking them onto separate streams measured the same: a
// long-horizon CTA leaves no shared memory for a second kernel's CTA on its SM).  Classes with fewer
// than 16 agents per SM are merged into the next larger class of the same solver family -- every launch ends with
// a tail in which the SMs run dry one by one, about as long as the longest agent's own SQP loop, and a class
// has to be worth that: on the 455 real scenarios (6050 agents, horizons 43..205) a threshold of 4 per SM gives
// 94.5 k QP/s, of 8 or more 120.8 k (one launch per solver family), no buckets 86.2 k (<= 96 steps: one-warp
// solver, above: CTA-wide solver; inside a family the arithmetic does not depend on the launch shape).
struct Bucket {
  int nt;      // longest horizon in the bucket
  int count, offset;   // segment of the grouped order
};

int horizon_class(int nt) { return launch_nt(nt); }   // (defined above run_refine)

// order: the agents to refine, in processing order (longest first); agent_nt: horizon of every agent of the batch.
// Returns the order grouped by bucket (largest horizons first) and the buckets.
void plan_buckets(const std::vector<int> &order, const std::vector<int> &agent_nt, int min_count,
                  std::vector<int> &grouped, std::vector<Bucket> &buckets) {
  const int n_cls = kMaxThreads / 16 + 1;
  std::vector<int> cnt(n_cls, 0), ntmax(n_cls, 0), target(n_cls);
  for (int a : order) {
    const int c = horizon_class(agent_nt[a]) / 16;
    cnt[c]++; ntmax[c] = std::max(ntmax[c], agent_nt[a]);
  }
  for (int c = 0; c < n_cls; ++c) target[c] = c;
  auto family = [](int c) { return 16 * c <= kOneWarpMaxNTHost ? 0 : 1; };
  auto next_used = [&](int c) { for (int d = c + 1; d < n_cls; ++d) if (cnt[d]) return d; return -1; };
  auto merge_up = [&](int c, int d) { cnt[d] += cnt[c]; ntmax[d] = std::max(ntmax[d], ntmax[c]); cnt[c] = 0; target[c] = d; };
  for (int c = 0; c < n_cls; ++c) {   // small classes move up inside their family
    if (!cnt[c] || cnt[c] >= min_count) continue;
    const int d = next_used(c);
    if (d >= 0 && family(d) == family(c)) merge_up(c, d);
  }
  for (;;) {                          // at most kMaxBuckets launches: fold the smallest class into its upper neighbour
    int used = 0, best = -1;
    for (int c = 0; c < n_cls; ++c) if (cnt[c]) { ++used; if (next_used(c) >= 0 && (best < 0 || cnt[c] < cnt[best])) best = c; }
    if (used <= kMaxBuckets || best < 0) break;
    merge_up(best, next_used(best));
  }
  auto final_of = [&](int c) { while (target[c] != c) c = target[c]; return c; };
  buckets.clear();
  std::vector<int> slot(n_cls, -1);
  for (int c = n_cls - 1; c >= 0; --c)
    if (cnt[c]) { slot[c] = (int)buckets.size(); buckets.push_back({ntmax[c], cnt[c], 0}); }
  int off = 0;
  for (auto &b : buckets) { b.offset = off; off += b.count; }
  grouped.resize(order.size());
  std::vector<int> fill(buckets.size(), 0);
  for (int a : order) {
    const int b = slot[final_of(horizon_class(agent_nt[a]) / 16)];
    grouped[buckets[b].offset + fill[b]++] = a;
  }
}

// Refine `order` (host copy of the processing order) bucket by bucket.  B.agent_order / n_active are replaced.
int run_refine_bucketed(csdo_handle *h, DevBatch B, const DevOut &O, const std::vector<int> &order,
                        const std::vector<int> &agent_nt, int max_k, cudaStream_t stream) {
  if (order.empty()) return CSDO_OK;
  std::vector<int> grouped;
  std::vector<Bucket> buckets;
  const char *off = getenv("CSDO_NO_BUCKETS");   // developer knob: one launch shaped for the longest horizon
  if (off && atoi(off)) {
    int nt = 0;
    for (int a : order) nt = std::max(nt, agent_nt[a]);
    grouped = order;
    buckets.push_back({nt, (int)order.size(), 0});
  } else {
    const char *mc = getenv("CSDO_BUCKET_MIN");   // developer knob: smallest class that keeps its own launch
    plan_buckets(order, agent_nt, mc ? std::max(1, atoi(mc)) : 16 * h->num_sms, grouped, buckets);
  }
  int rc;
  if ((rc = ensure(h, h->order_buf, grouped.size() * sizeof(int)))) return rc;
  if (set_err(h, "cudaMemcpyAsync H2D", cudaMemcpyAsync(h->order_buf.p, grouped.data(), grouped.size() * sizeof(int),
                                                        cudaMemcpyHostToDevice, stream)))
    return CSDO_ERR_CUDA;
  if (set_err(h, "init_outputs", launch_init_outputs(B, O, stream))) return CSDO_ERR_CUDA;
  const int nb = (int)buckets.size();
  for (int b = 0; b < nb; ++b) {
    B.agent_order = static_cast<const int *>(h->order_buf.p) + buckets[b].offset;
    B.n_active = buckets[b].count;
    if ((rc = run_refine(h, B, O, buckets[b].nt, max_k, stream, b, false, false, b == 0))) return rc;
  }
  h->buckets_used = nb;
  if (set_err(h, "aggregate_status", launch_aggregate_status(B, O, stream))) return CSDO_ERR_CUDA;
  h->last.launches += 2;
  return CSDO_OK;
}

}  // namespace

extern "C" {

void csdo_default_params(csdo_params *p) {
  // common/motion_planning.cc:54-109 on the shipped config.yaml; Constants are float
  const float r = 3.0f, deltat = 0.706f, LF = 2.0f, LB = 1.0f, carWidth = 2.0f, WB = 1.0f;
  std::memset(p, 0, sizeof(*p));
  p->f2x = (float)(1 / 4.0 * (3.0 * LF - LB));
  p->r2x = (float)(1 / 4.0 * (LF - 3.0 * LB));
  p->rv = (float)(1.0 / 2.0 * pow(pow(LF + LB, 2) / 4 + carWidth * carWidth, 0.5));
  p->WB = WB;
  p->steer_max = atan((double)WB / r);  // dsqp_solver.cc:1178
  p->LF = LF; p->LB = LB; p->car_width = carWidth;
  // sqp/utils.cc:34-59
  p->r_trust = 2.0; p->max_omega = 0.07; p->max_v = 1.0; p->delta_solution_threshold = 1.0;
  const int num_interpolation = 2;
  const double decelerate_factor = 0.8;
  p->dt = r * deltat / p->max_v / (num_interpolation + 1) / decelerate_factor;
  p->max_iter = 10; p->osqp_max_iter = 400; p->fixed_corridor = 0;
  // osqp_set_default_settings (OSQP 0.6.x) + the pinned rho interval
  p->adaptive_rho_interval = 25; p->scaling = 10; p->check_termination = 25; p->adaptive_rho = 1;
  p->rho = 0.1; p->sigma = 1e-6; p->alpha = 1.6;
  p->eps_abs = 1e-3; p->eps_rel = 1e-3; p->eps_prim_inf = 1e-4; p->eps_dual_inf = 1e-4;
  p->adaptive_rho_tolerance = 5.0;
  p->box_ds = 0.1; p->box_limit = 10.0;
}

const char *csdo_version(void) { return "csdo-dsqp-b200 0.1 sm_100a"; }

int csdo_create(const csdo_params *params, int device, csdo_handle **out) {
  if (!out) return CSDO_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return CSDO_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return CSDO_ERR_CUDA;
  if (prop.major != 10) return CSDO_ERR_CUDA;  // sm_100a code only
  DeviceGuard guard(device);
  csdo_handle *h = new csdo_handle();
  h->device = device;
  if (params) h->P = *params; else csdo_default_params(&h->P);
  if (h->P.osqp_max_iter < 1 || h->P.scaling < 0 || h->P.max_iter < 0) { delete h; return CSDO_ERR_INVALID; }
  h->num_sms = prop.multiProcessorCount;
  h->smem_limit = (int)prop.sharedMemPerBlockOptin;
  h->smem_limit_sm = (int)prop.sharedMemPerMultiprocessor;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return CSDO_ERR_CUDA; }
  *out = h;
  return CSDO_OK;
}

void csdo_destroy(csdo_handle *h) {
  if (!h) return;
  DeviceGuard guard(h->device);
  cudaStreamSynchronize(h->stream);
  for (auto &b : h->stage) if (b.p) cudaFree(b.p);
  if (h->scratch.p) cudaFree(h->scratch.p);
  if (h->queue.p) cudaFree(h->queue.p);
  if (h->step_cnt.p) cudaFree(h->step_cnt.p);
  if (h->pass_buf.p) cudaFree(h->pass_buf.p);
  if (h->tile_sum.p) cudaFree(h->tile_sum.p);
  if (h->order_buf.p) cudaFree(h->order_buf.p);
  cudaStreamDestroy(h->stream);
  delete h;
}

const char *csdo_last_error(const csdo_handle *h) { return h ? h->err.c_str() : "null handle"; }

int csdo_last_launch(const csdo_handle *h, csdo_launch_info *info) {
  if (!h || !info) return CSDO_ERR_INVALID;
  *info = h->last;
  return CSDO_OK;
}

int csdo_refine_device(csdo_handle *h, const csdo_batch *in, csdo_result *out, int max_nt, int max_planes,
                       void *cuda_stream) {
  if (!h || !in || !out) return CSDO_ERR_INVALID;
  if (!out->traj || !out->corridors || !out->status || !out->sqp_iters || !out->n_qp || !out->admm_iters ||
      !out->n_factor || !out->objective || !out->inst_status || !out->inst_static_legal) {
    h->err = "csdo_refine_device: every csdo_result array is required"; return CSDO_ERR_INVALID;
  }
  if (in->n_agents > 0 && (!in->inst_agent_ptr || !in->inst_nt || !in->inst_dims || !in->obs_ptr || !in->agent_off ||
                           !in->guess || !in->plane_ptr)) {
    h->err = "null array in batch"; return CSDO_ERR_INVALID;
  }
  DeviceGuard guard(h->device);
  DevBatch B = as_dev(in);
  DevOut O{out->traj, out->corridors, out->status, out->sqp_iters, out->n_qp, out->admm_iters,
           out->n_factor, out->objective, out->inst_status, out->inst_static_legal};
  cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->stream;
  h->last_stream = s;
  return run_refine(h, B, O, max_nt, max_planes, s);
}

int csdo_refine_device_hinted(csdo_handle *h, const csdo_batch *in, csdo_result *out, int max_planes,
                              const int32_t *host_inst_nt, const int32_t *host_inst_agent_ptr,
                              const int32_t *host_order, int32_t n_order, void *cuda_stream) {
  if (!h || !in || !out || !host_inst_nt || !host_inst_agent_ptr || n_order < 0 || (n_order > 0 && !host_order)) return CSDO_ERR_INVALID;
  if (!out->traj || !out->corridors || !out->status || !out->sqp_iters || !out->n_qp || !out->admm_iters ||
      !out->n_factor || !out->objective || !out->inst_status || !out->inst_static_legal) {
    h->err = "csdo_refine_device_hinted: every csdo_result array is required"; return CSDO_ERR_INVALID;
  }
  if (in->n_agents > 0 && (!in->inst_agent_ptr || !in->inst_nt || !in->inst_dims || !in->obs_ptr || !in->agent_off ||
                           !in->guess || !in->plane_ptr)) {
    h->err = "null array in batch"; return CSDO_ERR_INVALID;
  }
  if (in->n_agents == 0) return CSDO_OK;
  if (host_inst_agent_ptr[0] != 0 || host_inst_agent_ptr[in->n_inst] != in->n_agents) {
    h->err = "host_inst_agent_ptr does not cover the agents"; return CSDO_ERR_INVALID;
  }
  std::vector<int> agent_nt(in->n_agents);
  for (int i = 0; i < in->n_inst; ++i) {
    const int nt = host_inst_nt[i];
    if (nt < 3) { h->err = "horizon < 3"; return CSDO_ERR_INVALID; }
    if (nt > kMaxThreads) { h->err = "horizon exceeds 512 steps"; return CSDO_ERR_UNSUPPORTED; }
    if (host_inst_agent_ptr[i + 1] < host_inst_agent_ptr[i]) { h->err = "host_inst_agent_ptr not monotonic"; return CSDO_ERR_INVALID; }
    for (int a = host_inst_agent_ptr[i]; a < host_inst_agent_ptr[i + 1]; ++a) agent_nt[a] = nt;
  }
  std::vector<int> order;
  if (n_order > 0) {
    order.assign(host_order, host_order + n_order);
    std::vector<char> seen((size_t)in->n_agents, 0);
    for (int v : order) {
      if (v < 0 || v >= in->n_agents || seen[v]) { h->err = "host_order is not a list of distinct agent ids"; return CSDO_ERR_INVALID; }
      seen[v] = 1;
    }
  } else {   // every agent, longest horizon first
    order.resize(in->n_agents);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return agent_nt[x] > agent_nt[y]; });
  }
  DeviceGuard guard(h->device);
  DevBatch B = as_dev(in);
  DevOut O{out->traj, out->corridors, out->status, out->sqp_iters, out->n_qp, out->admm_iters,
           out->n_factor, out->objective, out->inst_status, out->inst_static_legal};
  cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->stream;
  h->last_stream = s;
  return run_refine_bucketed(h, B, O, order, agent_nt, max_planes, s);
}

int csdo_plan_horizon_buckets(int32_t n_agents, const int32_t *agent_nt, int32_t min_count, int32_t *order_out,
                              int32_t *bucket_nt, int32_t *bucket_count, int32_t max_buckets) {
  if (n_agents < 0 || (n_agents > 0 && (!agent_nt || !order_out)) || !bucket_nt || !bucket_count || max_buckets < 1) return -1;
  std::vector<int> nt(agent_nt, agent_nt + n_agents), order(n_agents), grouped;
  for (int a = 0; a < n_agents; ++a)
    if (nt[a] < 1 || nt[a] > kMaxThreads) return -1;
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return nt[x] > nt[y]; });
  std::vector<Bucket> buckets;
  if (n_agents) plan_buckets(order, nt, std::max(1, (int)min_count), grouped, buckets);
  if ((int)buckets.size() > max_buckets) return -1;
  for (int a = 0; a < n_agents; ++a) order_out[a] = grouped[a];
  for (size_t b = 0; b < buckets.size(); ++b) { bucket_nt[b] = buckets[b].nt; bucket_count[b] = buckets[b].count; }
  return (int)buckets.size();
}

int csdo_aggregate_status_device(csdo_handle *h, const csdo_batch *in, csdo_result *out, void *cuda_stream) {
  if (!h || !in || !out || !out->status || !out->inst_status || !in->inst_agent_ptr) return CSDO_ERR_INVALID;
  DeviceGuard guard(h->device);
  cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->stream;
  if (in->n_inst == 0) return CSDO_OK;
  DevBatch B = as_dev(in);
  DevOut O{out->traj, out->corridors, out->status, out->sqp_iters, out->n_qp, out->admm_iters,
           out->n_factor, out->objective, out->inst_status, out->inst_static_legal};
  if (set_err(h, "aggregate_status", launch_aggregate_status(B, O, s))) return CSDO_ERR_CUDA;
  return CSDO_OK;
}

int csdo_sync(csdo_handle *h) {
  if (!h) return CSDO_ERR_INVALID;
  DeviceGuard guard(h->device);
  cudaStream_t s = h->last_stream ? h->last_stream : h->stream;
  if (set_err(h, "csdo_sync", cudaStreamSynchronize(s))) return CSDO_ERR_CUDA;
  if (!h->queue.p) return CSDO_OK;
  int qerr = 0;
  for (int b = 0; b < h->buckets_used && h->queue.cap >= (size_t)(b + 1) * kQueueCtrlBytes; ++b) {
    int e = 0;
    if (set_err(h, "csdo_sync", cudaMemcpy(&e, reinterpret_cast<int *>(static_cast<char *>(h->queue.p) + (size_t)b * kQueueCtrlBytes) + kQError,
                                           sizeof(int), cudaMemcpyDeviceToHost)))
      return CSDO_ERR_CUDA;
    if (e) qerr = (qerr == 2 || e == 2) ? 2 : e;
  }
  if (qerr == 2) { h->err = "refine: an agent has more planes than max_planes"; return CSDO_ERR_INVALID; }
  if (qerr) { h->err = "refine: work queue stalled"; return CSDO_ERR_CUDA; }
  return CSDO_OK;
}

int csdo_refine(csdo_handle *h, const csdo_batch *in, csdo_result *out) {
  if (!h || !in || !out) return CSDO_ERR_INVALID;
  DeviceGuard guard(h->device);
  Meta m;
  int rc = validate_host(h, in, true, m);
  if (rc) return rc;
  if (in->n_agents == 0) return CSDO_OK;
  DevBatch B;
  if ((rc = upload_batch(h, in, m, true, B))) return rc;
  // processing order: longest problems first (unless the caller gave one)
  std::vector<int> order(in->n_agents);
  const int n_listed = in->n_active > 0 ? in->n_active : in->n_agents;
  if (in->agent_order) std::copy(in->agent_order, in->agent_order + n_listed, order.begin());
  else {
    std::iota(order.begin(), order.end(), 0);
    std::vector<int64_t> cost(in->n_agents);
    for (int a = 0; a < in->n_agents; ++a)
      cost[a] = 13 * (in->agent_off[a + 1] - in->agent_off[a]) + 4 * (int64_t)(in->plane_ptr[a + 1] - in->plane_ptr[a]);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cost[x] > cost[y]; });
  }
  order.resize(n_listed);
  std::vector<int> agent_nt(in->n_agents);
  for (int a = 0; a < in->n_agents; ++a) agent_nt[a] = (int)(in->agent_off[a + 1] - in->agent_off[a]);
  DevOut O;
  const size_t A = in->n_agents, I = in->n_inst;
  if ((rc = devalloc(h, 11, 6 * (size_t)m.steps, &O.traj))) return rc;
  if ((rc = devalloc(h, 12, 8 * (size_t)m.steps, &O.corridors))) return rc;
  int *ints = nullptr;
  if ((rc = devalloc(h, 13, 5 * A + 2 * I, &ints))) return rc;
  O.status = ints; O.sqp_iters = ints + A; O.n_qp = ints + 2 * A; O.admm_iters = ints + 3 * A;
  O.n_factor = ints + 4 * A; O.inst_status = ints + 5 * A; O.inst_static_legal = ints + 5 * A + I;
  if ((rc = devalloc(h, 14, A, &O.objective))) return rc;
  if ((rc = run_refine_bucketed(h, B, O, order, agent_nt, m.max_k, h->stream))) return rc;
  if ((rc = download(h, out->traj, O.traj, 6 * (size_t)m.steps))) return rc;
  if ((rc = download(h, out->corridors, O.corridors, 8 * (size_t)m.steps))) return rc;
  if ((rc = download(h, out->status, O.status, A))) return rc;
  if ((rc = download(h, out->sqp_iters, O.sqp_iters, A))) return rc;
  if ((rc = download(h, out->n_qp, O.n_qp, A))) return rc;
  if ((rc = download(h, out->admm_iters, O.admm_iters, A))) return rc;
  if ((rc = download(h, out->n_factor, O.n_factor, A))) return rc;
  if ((rc = download(h, out->objective, O.objective, A))) return rc;
  if ((rc = download(h, out->inst_status, O.inst_status, I))) return rc;
  if ((rc = download(h, out->inst_static_legal, O.inst_static_legal, I))) return rc;
  h->last_stream = h->stream;
  if ((rc = csdo_sync(h))) return rc;  // also reports the work queue's error flag
  if (getenv("CSDO_PROFILE")) {   // developer aid (needs a -DCSDO_DEV_TIMERS build): per-phase cycles, summed over CTAs
    unsigned long long ph[8];
    cudaMemcpy(ph, static_cast<char *>(h->queue.p) + 8, sizeof(ph), cudaMemcpyDeviceToHost);
    static const char *names[8] = {"corridor", "assemble", "ruiz", "factor", "solve", "rows", "check", "other"};
    double tot = 0;
    for (int k = 0; k < 8; ++k) tot += (double)ph[k];
    for (int k = 0; k < 8; ++k)
      fprintf(stderr, "[csdo profile] %-9s %6.2f%%  %.3e cycles\n", names[k], 100.0 * ph[k] / tot, (double)ph[k]);
    unsigned long long dbg[32];
    csdo::read_debug_counters(dbg);
    static const char *dn[12] = {"S1 sweeps", "S1 barrier", "S2+barrier", "BCR levels", "S4 sweeps", "S4 scatter+barrier",
                                 "-", "-", "rows fixed", "rows planes", "rows barrier", "rows finish"};
    for (int k = 0; k < 12; ++k)
      if (dbg[k]) fprintf(stderr, "[csdo profile] solve/rows part %-20s %.3e cycles\n", dn[k], (double)dbg[k]);
  }
  return CSDO_OK;
}

int csdo_corridors(csdo_handle *h, const csdo_batch *in, int double_centres, double *corridors,
                   int32_t *box_status, int32_t *inst_static_legal) {
  if (!h || !in || !corridors) return CSDO_ERR_INVALID;
  DeviceGuard guard(h->device);
  Meta m;
  int rc = validate_host(h, in, false, m);
  if (rc) return rc;
  if (in->n_agents == 0) return CSDO_OK;
  DevBatch B;
  if ((rc = upload_batch(h, in, m, false, B))) return rc;
  double *d_corr; int *d_bs, *d_legal;
  if ((rc = devalloc(h, 12, 8 * (size_t)m.steps, &d_corr))) return rc;
  if ((rc = devalloc(h, 13, 4 * (size_t)m.steps + in->n_inst, &d_bs))) return rc;
  d_legal = d_bs + 4 * (size_t)m.steps;
  if (set_err(h, "launch_corridors", launch_corridors(B, h->P, double_centres, d_corr, d_bs, d_legal, h->stream)))
    return CSDO_ERR_CUDA;
  h->last.launches = 2;
  if ((rc = download(h, corridors, d_corr, 8 * (size_t)m.steps))) return rc;
  if ((rc = download(h, box_status, d_bs, 4 * (size_t)m.steps))) return rc;
  if ((rc = download(h, inst_static_legal, d_legal, (size_t)in->n_inst))) return rc;
  if (set_err(h, "corridors", cudaStreamSynchronize(h->stream))) return CSDO_ERR_CUDA;
  return CSDO_OK;
}

int csdo_planes_count_device(csdo_handle *h, const csdo_batch *in, int64_t total_steps, int32_t *step_off,
                             int32_t *plane_ptr, int32_t *inst_inter_legal, int64_t *total_planes, void *cuda_stream) {
  if (!h || !in || !step_off || !plane_ptr || !inst_inter_legal) return CSDO_ERR_INVALID;
  DeviceGuard guard(h->device);
  cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->stream;
  if (total_planes) *total_planes = 0;
  if (in->n_agents == 0) return CSDO_OK;
  const DevBatch B = as_dev(in);
  const int n_tiles = (int)((total_steps + 2047) / 2048);
  int rc = ensure(h, h->tile_sum, ((size_t)n_tiles + 2) * sizeof(int));
  if (rc) return rc;
  if (set_err(h, "launch_planes_count", launch_planes_count(B, h->P, total_steps, step_off, inst_inter_legal, s))) return CSDO_ERR_CUDA;
  if (set_err(h, "launch_plane_offsets",
              launch_plane_offsets(B, total_steps, step_off, static_cast<int *>(h->tile_sum.p), plane_ptr, s)))
    return CSDO_ERR_CUDA;
  h->last.launches = 6;
  if (total_planes) {
    int tot = 0;
    if (set_err(h, "planes total", cudaMemcpyAsync(&tot, step_off + total_steps, sizeof(int), cudaMemcpyDeviceToHost, s)) ||
        set_err(h, "planes_count", cudaStreamSynchronize(s)))
      return CSDO_ERR_CUDA;
    if (tot < 0) { h->err = "plane count overflows int32"; return CSDO_ERR_UNSUPPORTED; }
    *total_planes = tot;
  }
  return CSDO_OK;
}

int csdo_planes_fill_device(csdo_handle *h, const csdo_batch *in, const int32_t *step_off, int32_t *plane_t,
                            double *plane_abc, int32_t *plane_partner, void *cuda_stream) {
  if (!h || !in || !step_off) return CSDO_ERR_INVALID;
  DeviceGuard guard(h->device);
  cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->stream;
  if (in->n_agents == 0) return CSDO_OK;
  if (!plane_t || !plane_abc) return CSDO_ERR_INVALID;
  if (set_err(h, "launch_planes_fill", launch_planes_fill(as_dev(in), h->P, step_off, plane_t, plane_abc, plane_partner, s)))
    return CSDO_ERR_CUDA;
  h->last.launches = 1;
  return CSDO_OK;
}

int csdo_planes_count(csdo_handle *h, const csdo_batch *in, int32_t *plane_ptr, int32_t *inst_inter_legal) {
  if (!h || !in || !plane_ptr) return CSDO_ERR_INVALID;
  DeviceGuard guard(h->device);
  Meta m;
  int rc = validate_host(h, in, false, m);
  if (rc) return rc;
  plane_ptr[0] = 0;
  if (inst_inter_legal) for (int i = 0; i < in->n_inst; ++i) inst_inter_legal[i] = 1;
  if (in->n_agents == 0) return CSDO_OK;
  DevBatch B;
  if ((rc = upload_batch(h, in, m, false, B))) return rc;
  if ((rc = ensure(h, h->step_cnt, ((size_t)m.steps + 1 + in->n_inst + in->n_agents + 1) * sizeof(int)))) return rc;
  int *d_off = static_cast<int *>(h->step_cnt.p), *d_legal = d_off + m.steps + 1, *d_ptr = d_legal + in->n_inst;
  csdo_batch dv = *in;
  dv.inst_agent_ptr = B.inst_agent_ptr; dv.inst_nt = B.inst_nt; dv.inst_dims = B.inst_dims; dv.obs_ptr = B.obs_ptr;
  dv.obs = B.obs; dv.agent_off = B.agent_off; dv.guess = B.guess; dv.plane_ptr = nullptr; dv.plane_t = nullptr;
  dv.plane_abc = nullptr; dv.agent_order = nullptr; dv.n_active = 0;
  int64_t total = 0;
  if ((rc = csdo_planes_count_device(h, &dv, m.steps, d_off, d_ptr, d_legal, &total, h->stream))) return rc;
  if ((rc = download(h, plane_ptr, d_ptr, (size_t)in->n_agents + 1))) return rc;
  if ((rc = download(h, inst_inter_legal, d_legal, (size_t)in->n_inst))) return rc;
  if (set_err(h, "planes_count", cudaStreamSynchronize(h->stream))) return CSDO_ERR_CUDA;
  return CSDO_OK;
}

int csdo_planes_fill_partners(csdo_handle *h, const csdo_batch *in, const int32_t *plane_ptr, int32_t *plane_t,
                              double *plane_abc, int32_t *plane_partner) {
  if (!h || !in || !plane_ptr) return CSDO_ERR_INVALID;
  DeviceGuard guard(h->device);
  Meta m;
  int rc = validate_host(h, in, false, m);
  if (rc) return rc;
  if (in->n_agents == 0) return CSDO_OK;
  const size_t total = (size_t)plane_ptr[in->n_agents];
  if (total == 0) return CSDO_OK;
  if (!plane_t || !plane_abc) return CSDO_ERR_INVALID;
  if (h->step_cnt.cap < ((size_t)m.steps + 1) * sizeof(int)) { h->err = "csdo_planes_count must precede csdo_planes_fill"; return CSDO_ERR_INVALID; }
  DevBatch B;
  if ((rc = upload_batch(h, in, m, false, B))) return rc;
  int *d_t, *d_partner = nullptr; double *d_abc;
  if ((rc = devalloc(h, 8, total, &d_t))) return rc;
  if ((rc = devalloc(h, 9, 12 * total, &d_abc))) return rc;
  if (plane_partner && (rc = devalloc(h, 15, total, &d_partner))) return rc;
  if (set_err(h, "launch_planes_fill",
              launch_planes_fill(B, h->P, static_cast<int *>(h->step_cnt.p), d_t, d_abc, d_partner, h->stream)))
    return CSDO_ERR_CUDA;
  h->last.launches = 1;
  if ((rc = download(h, plane_t, d_t, total))) return rc;
  if ((rc = download(h, plane_abc, d_abc, 12 * total))) return rc;
  if (plane_partner && (rc = download(h, plane_partner, d_partner, total))) return rc;
  if (set_err(h, "planes_fill", cudaStreamSynchronize(h->stream))) return CSDO_ERR_CUDA;
  return CSDO_OK;
}

int csdo_planes_fill(csdo_handle *h, const csdo_batch *in, const int32_t *plane_ptr, int32_t *plane_t,
                     double *plane_abc) {
  return csdo_planes_fill_partners(h, in, plane_ptr, plane_t, plane_abc, nullptr);
}

int csdo_planes_from_pairs(csdo_handle *h, const csdo_batch *in, int64_t n_pairs, const int32_t *pairs,
                           int32_t *plane_ptr, int32_t *plane_t, double *plane_abc) {
  if (!h || !in || !plane_ptr || n_pairs < 0 || (n_pairs > 0 && !pairs)) return CSDO_ERR_INVALID;
  DeviceGuard guard(h->device);
  Meta m;
  int rc = validate_host(h, in, false, m);
  if (rc) return rc;
  const int A = in->n_agents;
  for (int a = 0; a <= A; ++a) plane_ptr[a] = 0;
  if (A == 0 || n_pairs == 0) return CSDO_OK;
  if (2 * n_pairs > INT32_MAX) { h->err = "plane count overflows int32"; return CSDO_ERR_UNSUPPORTED; }
  if (!plane_t || !plane_abc) return CSDO_ERR_INVALID;
  // every pair pushes one plane to agent i and one to agent j, in list order (inter_agent_cons.cc:136-137)
  for (int64_t p = 0; p < n_pairs; ++p) {
    const int t = pairs[3 * p], i = pairs[3 * p + 1], j = pairs[3 * p + 2];
    if (i < 0 || j < 0 || i >= A || j >= A || i == j || t < 0 || t >= in->agent_off[i + 1] - in->agent_off[i] ||
        in->agent_off[j + 1] - in->agent_off[j] != in->agent_off[i + 1] - in->agent_off[i]) {
      h->err = "invalid neighbour pair"; return CSDO_ERR_INVALID;
    }
    plane_ptr[i + 1]++; plane_ptr[j + 1]++;
  }
  for (int a = 0; a < A; ++a) plane_ptr[a + 1] += plane_ptr[a];
  std::vector<int> cursor(plane_ptr, plane_ptr + A), pos((size_t)2 * n_pairs);
  for (int64_t p = 0; p < n_pairs; ++p) {
    pos[2 * p] = cursor[pairs[3 * p + 1]]++;
    pos[2 * p + 1] = cursor[pairs[3 * p + 2]]++;
  }
  DevBatch B;
  if ((rc = upload_batch(h, in, m, false, B))) return rc;
  const int *d_pairs, *d_pos;
  if ((rc = upload(h, 16, pairs, (size_t)3 * n_pairs, &d_pairs))) return rc;
  if ((rc = upload(h, 17, pos.data(), pos.size(), &d_pos))) return rc;
  int *d_t; double *d_abc;
  const size_t total = (size_t)2 * n_pairs;
  if ((rc = devalloc(h, 8, total, &d_t))) return rc;
  if ((rc = devalloc(h, 9, 12 * total, &d_abc))) return rc;
  if (set_err(h, "launch_planes_from_pairs",
              launch_planes_from_pairs(B, h->P, n_pairs, d_pairs, d_pos, d_t, d_abc, h->stream)))
    return CSDO_ERR_CUDA;
  h->last.launches = 1;
  if ((rc = download(h, plane_t, d_t, total))) return rc;
  if ((rc = download(h, plane_abc, d_abc, 12 * total))) return rc;
  if (set_err(h, "planes_from_pairs", cudaStreamSynchronize(h->stream))) return CSDO_ERR_CUDA;
  return CSDO_OK;
}

}  // extern "C"
