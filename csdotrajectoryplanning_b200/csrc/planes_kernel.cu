// Pre-process kernels: neighbour pairs and separating planes.
//   findNeighborPairsByTrustRegion   sqp/inter_agent_cons.cc:12-49
//   calcPerpendicular                sqp/inter_agent_cons.cc:54-69
//   calcEqualInterPlanes             sqp/inter_agent_cons.cc:71-140
//   State (float disc centres, agentDistance, agentCollision)
//                                    common/motion_planning.h:113-217
// One thread per (agent, time step); partners are scanned in ascending agent
// order, which reproduces the reference's per-agent push order (sorted by t,
// then by partner) without atomics.  Compiled with -fmad=false: given the same
// disc centres the planes are bit-identical to the reference arithmetic.
#include <math.h>

#include "dsqp_launch.h"

namespace csdo {

struct DState {
  double yaw;
  float xc, yc, xf, xr, yf, yr;
};

__device__ __forceinline__ DState make_state(const csdo_params &P, double x, double y, double yaw) {
  DState s;
  const double cs = cos(yaw), sn = sin(yaw);
  s.yaw = yaw;
  s.xf = (float)(x + P.f2x * cs); s.xr = (float)(x + P.r2x * cs);
  s.yf = (float)(y + P.f2x * sn); s.yr = (float)(y + P.r2x * sn);
  const float LF = (float)P.LF, LB = (float)P.LB;
  const float d = (LF + LB) / 2 - LB;
  s.xc = (float)(x + d * cs);
  s.yc = (float)(y + d * sn);
  return s;
}

// State::agentDistance: float differences, squared and summed in double
__device__ __forceinline__ double agent_distance(const DState &a, const DState &b) {
  double d, e, p, q;
  p = (double)(a.xf - b.xf); q = (double)(a.yf - b.yf); d = p * p + q * q;
  p = (double)(a.xf - b.xr); q = (double)(a.yf - b.yr); e = p * p + q * q; d = e < d ? e : d;
  p = (double)(a.xr - b.xf); q = (double)(a.yr - b.yf); e = p * p + q * q; d = e < d ? e : d;
  p = (double)(a.xr - b.xr); q = (double)(a.yr - b.yr); e = p * p + q * q; d = e < d ? e : d;
  return sqrt(d);
}

// State::agentCollision (PRCISE_COLLISION branch): separating axes, all float
__device__ __forceinline__ bool agent_collision(const csdo_params &P, const DState &a, const DState &o) {
  const float length = (float)P.LF + (float)P.LB, width = (float)P.car_width;
  const float shift_x = o.xc - a.xc, shift_y = o.yc - a.yc;
  const float cos_v = (float)cos(a.yaw), sin_v = (float)sin(a.yaw);
  const float cos_o = (float)cos(o.yaw), sin_o = (float)sin(o.yaw);
  const float half_l = length / 2, half_w = width / 2;
  const float dx1 = cos_v * length / 2, dy1 = sin_v * length / 2;
  const float dx2 = sin_v * width / 2, dy2 = -cos_v * width / 2;
  const float dx3 = cos_o * length / 2, dy3 = sin_o * length / 2;
  const float dx4 = sin_o * width / 2, dy4 = -cos_o * width / 2;
  return ((fabsf(shift_x * cos_v + shift_y * sin_v) <=
           fabsf(dx3 * cos_v + dy3 * sin_v) + fabsf(dx4 * cos_v + dy4 * sin_v) + half_l) &&
          (fabsf(shift_x * sin_v - shift_y * cos_v) <=
           fabsf(dx3 * sin_v - dy3 * cos_v) + fabsf(dx4 * sin_v - dy4 * cos_v) + half_w) &&
          (fabsf(shift_x * cos_o + shift_y * sin_o) <=
           fabsf(dx1 * cos_o + dy1 * sin_o) + fabsf(dx2 * cos_o + dy2 * sin_o) + half_l) &&
          (fabsf(shift_x * sin_o - shift_y * cos_o) <=
           fabsf(dx1 * sin_o - dy1 * cos_o) + fabsf(dx2 * sin_o - dy2 * cos_o) + half_w));
}

__device__ __forceinline__ void perpendicular(double rv, double x1, double y1, double x2, double y2,
                                              double &a, double &b, double &c1, double &c2) {
  a = x2 - x1;
  b = y2 - y1;
  const double c = (x1 * x1 + y1 * y1 - x2 * x2 - y2 * y2) / 2;
  const double d = sqrt((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2));
  c1 = c + rv * d;
  c2 = c - rv * d;
}

__device__ __forceinline__ int find_instance(const DevBatch &B, int a) {
  int lo = 0, hi = B.n_inst;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (B.inst_agent_ptr[mid] <= a) lo = mid; else hi = mid;
  }
  return lo;
}

// the two planes of one neighbour pair (calcEqualInterPlanes :71-140): plane_i for agent i, plane_j for j
__device__ __forceinline__ void pair_planes(const csdo_params &P, const DState &si, const DState &sj, double *oi,
                                            double *oj) {
  double a_f2f, b_f2f, c_f2f, c_f2f_, a_f2r, b_f2r, c_f2r, c_f2r_;
  double a_r2f, b_r2f, c_r2f, c_r2f_, a_r2r, b_r2r, c_r2r, c_r2r_;
  perpendicular(P.rv, si.xf, si.yf, sj.xf, sj.yf, a_f2f, b_f2f, c_f2f, c_f2f_);
  perpendicular(P.rv, si.xf, si.yf, sj.xr, sj.yr, a_f2r, b_f2r, c_f2r, c_f2r_);
  perpendicular(P.rv, si.xr, si.yr, sj.xf, sj.yf, a_r2f, b_r2f, c_r2f, c_r2f_);
  perpendicular(P.rv, si.xr, si.yr, sj.xr, sj.yr, a_r2r, b_r2r, c_r2r, c_r2r_);
  if (oi) {
    oi[0] = a_f2f; oi[1] = b_f2f; oi[2] = c_f2f; oi[3] = a_f2r; oi[4] = b_f2r; oi[5] = c_f2r;
    oi[6] = a_r2f; oi[7] = b_r2f; oi[8] = c_r2f; oi[9] = a_r2r; oi[10] = b_r2r; oi[11] = c_r2r;
  }
  if (oj) {  // its f2r is i's r2f (:131-135)
    oj[0] = -a_f2f; oj[1] = -b_f2f; oj[2] = -c_f2f_; oj[3] = -a_r2f; oj[4] = -b_r2f; oj[5] = -c_r2f_;
    oj[6] = -a_f2r; oj[7] = -b_f2r; oj[8] = -c_f2r_; oj[9] = -a_r2r; oj[10] = -b_r2r; oj[11] = -c_r2r_;
  }
}

template <bool FILL>
__global__ void planes_kernel(const DevBatch B, const csdo_params P, int *step_cnt, int *inst_inter_legal,
                              const int *step_off, int *plane_t, double *plane_abc, int *plane_partner) {
  const int a = (B.n_active > 0 && B.agent_order) ? B.agent_order[blockIdx.x] : (int)blockIdx.x;  // agent subset
  const int inst = find_instance(B, a);
  const int Nt = B.inst_nt[inst];
  const int a0 = B.inst_agent_ptr[inst], a1 = B.inst_agent_ptr[inst + 1];
  const int64_t off = B.agent_off[a];
  const double *g = B.guess + 6 * off;
  const double thr = 2 * sqrt(2.0) * P.r_trust;  // inter_agent_cons.cc:35
  for (int t = threadIdx.x; t < Nt; t += blockDim.x) {
    const DState sa = make_state(P, g[t], g[Nt + t], g[2 * Nt + t]);
    int cnt = 0;
    int k = FILL ? step_off[off + t] : 0;
    for (int b = a0; b < a1; ++b) {
      if (b == a) continue;
      const double *gb = B.guess + 6 * B.agent_off[b];
      const DState sb = make_state(P, gb[t], gb[Nt + t], gb[2 * Nt + t]);
      // the reference evaluates si.agentDistance(sj) with i < j
      const DState &si = a < b ? sa : sb;
      const DState &sj = a < b ? sb : sa;
      const double d = agent_distance(si, sj);
      if (!(d < thr)) continue;
      cnt++;
      if (!FILL) {
        if (a < b && agent_collision(P, si, sj)) atomicAnd(&inst_inter_legal[inst], 0);
        continue;
      }
      double *o = plane_abc + (size_t)12 * k;
      plane_t[k] = t;
      if (plane_partner) plane_partner[k] = b;
      pair_planes(P, si, sj, a < b ? o : nullptr, a < b ? nullptr : o);
      ++k;
    }
    if (!FILL) step_cnt[off + t] = cnt;
  }
}

__global__ void fill_int_kernel2(int *p, int n, int v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

cudaError_t launch_planes_count(const DevBatch &B, const csdo_params &P, int64_t total_steps, int *step_cnt,
                                int *inst_inter_legal, cudaStream_t stream) {
  fill_int_kernel2<<<(B.n_inst + 255) / 256, 256, 0, stream>>>(inst_inter_legal, B.n_inst, 1);
  const bool subset = B.n_active > 0 && B.agent_order;
  if (subset) cudaMemsetAsync(step_cnt, 0, (size_t)total_steps * sizeof(int), stream);  // the other agents: no planes
  if (B.n_agents > 0)
    planes_kernel<false><<<subset ? B.n_active : B.n_agents, 128, 0, stream>>>(B, P, step_cnt, inst_inter_legal, nullptr, nullptr, nullptr, nullptr);
  return cudaGetLastError();
}

cudaError_t launch_planes_fill(const DevBatch &B, const csdo_params &P, const int *step_off, int *plane_t,
                               double *plane_abc, int *plane_partner, cudaStream_t stream) {
  if (B.n_agents > 0)
    planes_kernel<true><<<(B.n_active > 0 && B.agent_order) ? B.n_active : B.n_agents, 128, 0, stream>>>(B, P, nullptr, nullptr, step_off, plane_t, plane_abc,
                                                         plane_partner);
  return cudaGetLastError();
}

// calcEqualInterPlanes for an explicit pair list: one thread per pair (t, i, j), i < j global agent ids;
// pos[2 p], pos[2 p + 1] = plane slots of agent i / agent j (the caller's push order)
__global__ void planes_from_pairs_kernel(const DevBatch B, const csdo_params P, int64_t n_pairs, const int *pairs,
                                         const int *pos, int *plane_t, double *plane_abc) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  const int t = pairs[3 * p], i = pairs[3 * p + 1], j = pairs[3 * p + 2];
  const int Nt = (int)(B.agent_off[i + 1] - B.agent_off[i]);
  const double *gi = B.guess + 6 * B.agent_off[i], *gj = B.guess + 6 * B.agent_off[j];
  const DState si = make_state(P, gi[t], gi[Nt + t], gi[2 * Nt + t]);
  const DState sj = make_state(P, gj[t], gj[Nt + t], gj[2 * Nt + t]);
  const int ki = pos[2 * p], kj = pos[2 * p + 1];
  plane_t[ki] = t;
  plane_t[kj] = t;
  pair_planes(P, si, sj, plane_abc + (size_t)12 * ki, plane_abc + (size_t)12 * kj);
}

cudaError_t launch_planes_from_pairs(const DevBatch &B, const csdo_params &P, int64_t n_pairs, const int *pairs,
                                     const int *pos, int *plane_t, double *plane_abc, cudaStream_t stream) {
  if (n_pairs > 0)
    planes_from_pairs_kernel<<<(unsigned)((n_pairs + 127) / 128), 128, 0, stream>>>(B, P, n_pairs, pairs, pos, plane_t,
                                                                                    plane_abc);
  return cudaGetLastError();
}

// ---- exclusive scan of the per-step plane counts (device resident, no host round trip) ----
constexpr int kScanTile = 2048;
__global__ void scan_tiles_kernel(int *v, int64_t n, int *tile_sum) {
  __shared__ int ws[8];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * 8;
  int x[8], s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = base + i < n ? v[base + i] : 0; s += x[i]; }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
  if (lane == 31) ws[warp] = inc;
  __syncthreads();
  int woff = 0;
  for (int w = 0; w < warp; ++w) woff += ws[w];
  int run = woff + inc - s;
#pragma unroll
  for (int i = 0; i < 8; ++i) { if (base + i < n) v[base + i] = run; run += x[i]; }
  if (threadIdx.x == 255) tile_sum[blockIdx.x] = run;
}
__global__ void scan_sums_kernel(int *tile_sum, int n_tiles, int *total) {  // one CTA
  __shared__ int ws[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < n_tiles; b0 += blockDim.x) {
    const int i = b0 + threadIdx.x;
    const int x = i < n_tiles ? tile_sum[i] : 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
    if (lane == 31) ws[warp] = inc;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; ++w) woff += ws[w];
    const int c0 = carry;
    if (i < n_tiles) tile_sum[i] = c0 + woff + inc - x;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = c0 + woff + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}
__global__ void scan_add_kernel(int *v, int64_t n, const int *tile_sum, const int *total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] += tile_sum[i / kScanTile];
  else if (i == n) v[n] = *total;
}
__global__ void plane_ptr_kernel(const int64_t *agent_off, int n_agents, const int *step_off, int *plane_ptr) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a <= n_agents) plane_ptr[a] = step_off[agent_off[a]];
}

// step_cnt [steps + 1] -> exclusive offsets in place (step_cnt[steps] = total); plane_ptr [n_agents + 1];
// tile_sum: ceil(steps / 2048) + 1 ints of scratch (the last one receives the total)
cudaError_t launch_plane_offsets(const DevBatch &B, int64_t steps, int *step_cnt, int *tile_sum, int *plane_ptr,
                                 cudaStream_t stream) {
  const int n_tiles = (int)((steps + kScanTile - 1) / kScanTile);
  if (n_tiles > 0) scan_tiles_kernel<<<n_tiles, 256, 0, stream>>>(step_cnt, steps, tile_sum);
  scan_sums_kernel<<<1, 1024, 0, stream>>>(tile_sum, n_tiles, tile_sum + n_tiles);
  scan_add_kernel<<<(unsigned)((steps + 1 + 255) / 256), 256, 0, stream>>>(step_cnt, steps, tile_sum, tile_sum + n_tiles);
  plane_ptr_kernel<<<(B.n_agents + 1 + 255) / 256, 256, 0, stream>>>(B.agent_off, B.n_agents, step_cnt, plane_ptr);
  return cudaGetLastError();
}

}  // namespace csdo
