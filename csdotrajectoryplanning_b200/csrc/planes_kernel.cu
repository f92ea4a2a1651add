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

template <bool FILL>
__global__ void planes_kernel(const DevBatch B, const csdo_params P, int *step_cnt, int *inst_inter_legal,
                              const int *step_off, int *plane_t, double *plane_abc) {
  const int a = blockIdx.x;
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
      double a_f2f, b_f2f, c_f2f, c_f2f_, a_f2r, b_f2r, c_f2r, c_f2r_;
      double a_r2f, b_r2f, c_r2f, c_r2f_, a_r2r, b_r2r, c_r2r, c_r2r_;
      perpendicular(P.rv, si.xf, si.yf, sj.xf, sj.yf, a_f2f, b_f2f, c_f2f, c_f2f_);
      perpendicular(P.rv, si.xf, si.yf, sj.xr, sj.yr, a_f2r, b_f2r, c_f2r, c_f2r_);
      perpendicular(P.rv, si.xr, si.yr, sj.xf, sj.yf, a_r2f, b_r2f, c_r2f, c_r2f_);
      perpendicular(P.rv, si.xr, si.yr, sj.xr, sj.yr, a_r2r, b_r2r, c_r2r, c_r2r_);
      double *o = plane_abc + (size_t)12 * k;
      plane_t[k] = t;
      if (a < b) {  // plane_i
        o[0] = a_f2f; o[1] = b_f2f; o[2] = c_f2f; o[3] = a_f2r; o[4] = b_f2r; o[5] = c_f2r;
        o[6] = a_r2f; o[7] = b_r2f; o[8] = c_r2f; o[9] = a_r2r; o[10] = b_r2r; o[11] = c_r2r;
      } else {      // plane_j: its f2r is i's r2f (:131-135)
        o[0] = -a_f2f; o[1] = -b_f2f; o[2] = -c_f2f_; o[3] = -a_r2f; o[4] = -b_r2f; o[5] = -c_r2f_;
        o[6] = -a_f2r; o[7] = -b_f2r; o[8] = -c_f2r_; o[9] = -a_r2r; o[10] = -b_r2r; o[11] = -c_r2r_;
      }
      ++k;
    }
    if (!FILL) step_cnt[off + t] = cnt;
  }
}

__global__ void fill_int_kernel2(int *p, int n, int v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

cudaError_t launch_planes_count(const DevBatch &B, const csdo_params &P, int *step_cnt, int *inst_inter_legal,
                                cudaStream_t stream) {
  fill_int_kernel2<<<(B.n_inst + 255) / 256, 256, 0, stream>>>(inst_inter_legal, B.n_inst, 1);
  if (B.n_agents > 0)
    planes_kernel<false><<<B.n_agents, 128, 0, stream>>>(B, P, step_cnt, inst_inter_legal, nullptr, nullptr, nullptr);
  return cudaGetLastError();
}

cudaError_t launch_planes_fill(const DevBatch &B, const csdo_params &P, const int *step_off, int *plane_t,
                               double *plane_abc, cudaStream_t stream) {
  if (B.n_agents > 0)
    planes_kernel<true><<<B.n_agents, 128, 0, stream>>>(B, P, nullptr, nullptr, step_off, plane_t, plane_abc);
  return cudaGetLastError();
}

}  // namespace csdo
