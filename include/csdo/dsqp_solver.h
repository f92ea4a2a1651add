// Drop-in C++ interface of the DSQP refine stage on top of the C ABI
// (include/csdo_dsqp.h).  Header-only; link with libcsdo_dsqp.so.
//
// Two modes.
//  * INSIDE THE REFERENCE TREE (define CSDO_WITH_REFERENCE_TYPES before the include, with the reference
//    root on the include path): the header uses the reference's OWN types --
//    libMultiRobotPlanning::{OptimizeResult, QpParm, Corridor, InterPlane} and the global Location with
//    its std::hash (common/motion_planning.h:80-109, sqp/common.h:14-52, sqp/corridor.h:8-11,
//    sqp/inter_agent_cons.h:47-63) -- and adds only
//        class SolverDSQP                                   sqp/dsqp_solver.h:24-47 (constructor = refine)
//        csdo_b200::findNeighborPairsByTrustRegion          sqp/inter_agent_cons.h:40-45 (same arguments)
//        csdo_b200::calcEqualInterPlanes                    sqp/inter_agent_cons.h:70-73 (same arguments)
//        csdo_b200::buildInterPlanes                        both in one call
//    so csdo.cc compiles with `#include "sqp/dsqp_solver.h"` replaced by this header (INTEGRATION.md shows
//    the diff; tests/cpp/ref_tree_compile.cpp is csdo.cc:111-167 built that way).  InterpolateInitalGuess
//    and dumpSolutions stay the reference's own (sqp/inter_agent_cons.cc needs neither Eigen nor OSQP).
//  * STANDALONE (default): the same type names with the same fields are defined here, Location gets
//    operator< / operator== and std::hash<Location>, and the two plane functions are also visible in
//    namespace libMultiRobotPlanning under the reference's names.
// Differences a maintainer has to know are listed in INTEGRATION.md: no Eigen/OSQP types (solveOSQP is
// gone), obstacles may be any iterable of Location (std::unordered_set<Location> included), and one
// extra optional constructor argument carries the vehicle constants that the reference keeps in the
// global `Constants` class.
#pragma once

#include <algorithm>
#include <array>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <stdexcept>
#include <string>
#include <tuple>
#include <unordered_set>
#include <vector>

#include "../csdo_dsqp.h"

#if defined(CSDO_WITH_REFERENCE_TYPES)
#include "common/motion_planning.h"
#include "sqp/common.h"
#include "sqp/corridor.h"  // pulls sqp/inter_agent_cons.h (InterPlane, InterpolateInitalGuess, dumpSolutions)
#else

namespace libMultiRobotPlanning {

struct OptimizeResult {  // sqp/common.h:14-22
  double x = 0, y = 0, yaw = 0, v = 0, a = 0, steer = 0, d_steer = 0;
};

struct QpParm {  // sqp/common.h:39-52
  double r_trust = 2.0;
  double max_omega = 0.07;
  double max_v = 1.0;
  double max_iter = 10;
  double delta_solution_threshold = 1.0;
  double max_violation = 0.001;
  int osqp_max_iter = 400;
  double dt = 0.0;
  int num_interpolation = 2;
  bool fixed_corridor = false;
};

struct Corridor {  // sqp/corridor.h:8-11
  double xf_min, xf_max, yf_min, yf_max;
  double xr_min, xr_max, yr_min, yr_max;
};

struct InterPlane {  // sqp/inter_agent_cons.h:47-63
  int t;
  double a_f2f, b_f2f, c_f2f, a_f2r, b_f2r, c_f2r;
  double a_r2f, b_r2f, c_r2f, a_r2r, b_r2r, c_r2r;
};

}  // namespace libMultiRobotPlanning

struct Location {  // common/motion_planning.h:80-97
  Location(double x, double y, double r) : x(x), y(y), r(r) {}
  double x, y, r;
  bool operator<(const Location &o) const { return std::tie(x, y, r) < std::tie(o.x, o.y, o.r); }
  bool operator==(const Location &o) const { return std::tie(x, y, r) == std::tie(o.x, o.y, o.r); }
};

namespace std {
template <>
struct hash<Location> {  // common/motion_planning.h:99-109 hashes (x, y); any mix of the two serves here
  size_t operator()(const Location &s) const {
    size_t seed = std::hash<double>()(s.x);
    seed ^= std::hash<double>()(s.y) + 0x9e3779b97f4a7c15ull + (seed << 6) + (seed >> 2);
    return seed;
  }
};
}  // namespace std

#endif  // CSDO_WITH_REFERENCE_TYPES

namespace csdo_detail {

using libMultiRobotPlanning::InterPlane;
using libMultiRobotPlanning::OptimizeResult;

inline void check(int rc, csdo_handle *h, const char *what) {
  if (rc != CSDO_OK) throw std::runtime_error(std::string(what) + ": " + (h ? csdo_last_error(h) : "no handle"));
}

// A library handle (stream, scratch, work queue) is kept per (device, parameters) for the life of the process
// instead of being created and destroyed by every call: a fresh handle costs ~4-10 ms (stream + first
// allocations) next to ~16 ms for refining one 25-agent instance.  Calls that share a handle are serialised.
class HandleLease {
 public:
  HandleLease(const csdo_params &P, int device) : lock_(mutex()) {
    for (Entry &e : cache())
      if (e.device == device && std::memcmp(&e.params, &P, sizeof(csdo_params)) == 0) { h_ = e.h; return; }
    csdo_handle *h = nullptr;
    check(csdo_create(&P, device, &h), nullptr, "csdo_create");
    cache().push_back(Entry{device, P, h});
    h_ = h;
  }
  csdo_handle *get() const { return h_; }

 private:
  struct Entry { int device; csdo_params params; csdo_handle *h; };
  // (never destroyed: at static-destruction time the CUDA runtime may already be gone; the driver reclaims
  // the stream and the allocations when the process ends)
  static std::vector<Entry> &cache() { static std::vector<Entry> *v = new std::vector<Entry>(); return *v; }
  static std::mutex &mutex() { static std::mutex m; return m; }
  std::unique_lock<std::mutex> lock_;
  csdo_handle *h_ = nullptr;
};

// one instance -> the flat batch layout of csdo_dsqp.h
struct Packed {
  int32_t inst_agent_ptr[2] = {0, 0};
  int32_t inst_nt[1] = {0};
  double dims[2] = {0, 0};
  int32_t obs_ptr[2] = {0, 0};
  std::vector<double> obs, guess, plane_abc;
  std::vector<int64_t> agent_off;
  std::vector<int32_t> plane_ptr, plane_t;
  csdo_batch view() const {
    csdo_batch b{};
    b.n_inst = 1; b.n_agents = inst_agent_ptr[1];
    b.inst_agent_ptr = inst_agent_ptr; b.inst_nt = inst_nt; b.inst_dims = dims;
    b.obs_ptr = obs_ptr; b.obs = obs.data(); b.agent_off = agent_off.data(); b.guess = guess.data();
    b.plane_ptr = plane_ptr.data(); b.plane_t = plane_t.data(); b.plane_abc = plane_abc.data();
    b.agent_order = nullptr;
    return b;
  }
};

inline void pack_guess(const std::vector<std::vector<OptimizeResult>> &x0_bar, Packed &p) {
  const int Na = (int)x0_bar.size(), Nt = Na ? (int)x0_bar[0].size() : 0;
  p.inst_agent_ptr[1] = Na; p.inst_nt[0] = Nt;
  p.agent_off.resize(Na + 1);
  p.guess.assign((size_t)6 * Na * Nt, 0.0);
  for (int a = 0; a <= Na; ++a) p.agent_off[a] = (int64_t)a * Nt;
  for (int a = 0; a < Na; ++a) {
    if ((int)x0_bar[a].size() != Nt) throw std::runtime_error("all agents must share the horizon");
    double *g = p.guess.data() + (size_t)6 * a * Nt;
    for (int t = 0; t < Nt; ++t) {
      const OptimizeResult &r = x0_bar[a][t];
      g[t] = r.x; g[Nt + t] = r.y; g[2 * Nt + t] = r.yaw; g[3 * Nt + t] = r.steer;
      g[4 * Nt + t] = r.v; g[5 * Nt + t] = r.d_steer;
    }
  }
  p.plane_ptr.assign(Na + 1, 0);
}

}  // namespace csdo_detail

namespace csdo_b200 {

using libMultiRobotPlanning::InterPlane;
using libMultiRobotPlanning::OptimizeResult;

namespace detail {
inline csdo_params plane_params(const csdo_params *params, const double *r_trust, const double *rv) {
  csdo_params P;
  if (params) P = *params; else csdo_default_params(&P);
  if (r_trust) P.r_trust = *r_trust;
  if (rv) P.rv = *rv;
  return P;
}
inline void unpack_planes(const csdo_detail::Packed &p, std::vector<std::vector<InterPlane>> &inter_planes) {
  const int Na = p.inst_agent_ptr[1];
  inter_planes.assign(Na, {});
  for (int a = 0; a < Na; ++a)
    for (int k = p.plane_ptr[a]; k < p.plane_ptr[a + 1]; ++k) {
      const double *q = p.plane_abc.data() + (size_t)12 * k;
      inter_planes[a].push_back(InterPlane{p.plane_t[k], q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7], q[8], q[9],
                                           q[10], q[11]});
    }
}
// count + fill through the C ABI; partner (optional) receives the other agent of every plane
inline bool build(const std::vector<std::vector<OptimizeResult>> &x0_bar, const csdo_params &P, int device,
                  csdo_detail::Packed &p, std::vector<int32_t> *partner) {
  csdo_detail::pack_guess(x0_bar, p);
  csdo_detail::HandleLease lease(P, device);
  csdo_handle *h = lease.get();
  const int Na = p.inst_agent_ptr[1];
  int32_t legal = 1;
  csdo_batch b = p.view();
  csdo_detail::check(csdo_planes_count(h, &b, p.plane_ptr.data(), &legal), h, "csdo_planes_count");
  const int total = p.plane_ptr[Na];
  p.plane_t.assign(total, 0);
  p.plane_abc.assign((size_t)12 * total, 0.0);
  if (partner) partner->assign(total, 0);
  csdo_detail::check(csdo_planes_fill_partners(h, &b, p.plane_ptr.data(), p.plane_t.data(), p.plane_abc.data(),
                                               partner ? partner->data() : nullptr),
                     h, "csdo_planes_fill");
  return legal != 0;
}
}  // namespace detail

// findNeighborPairsByTrustRegion (sqp/inter_agent_cons.cc:12-49), same arguments and result: the list of
// (t, i, j), i < j, in (t, i, j)-lexicographic order, and initial_inter_legal.
inline bool findNeighborPairsByTrustRegion(const std::vector<std::vector<OptimizeResult>> &solution, const double &r,
                                           const double &rv, std::vector<std::array<int, 3>> &neighbor_pairs,
                                           const csdo_params *vehicle = nullptr, int device = 0) {
  csdo_detail::Packed p;
  std::vector<int32_t> partner;
  const bool legal = detail::build(solution, detail::plane_params(vehicle, &r, &rv), device, p, &partner);
  neighbor_pairs.clear();
  const int Na = p.inst_agent_ptr[1];
  for (int a = 0; a < Na; ++a)
    for (int k = p.plane_ptr[a]; k < p.plane_ptr[a + 1]; ++k)
      if (a < partner[k]) neighbor_pairs.push_back(std::array<int, 3>{p.plane_t[k], a, partner[k]});
  std::sort(neighbor_pairs.begin(), neighbor_pairs.end());
  return legal;
}

// calcEqualInterPlanes (sqp/inter_agent_cons.cc:71-140), same arguments: the planes of the listed pairs,
// pushed to both agents in list order.
inline void calcEqualInterPlanes(const std::vector<std::vector<OptimizeResult>> &x0_bar,
                                 const std::vector<std::array<int, 3>> &neighbor_pairs,
                                 std::vector<std::vector<InterPlane>> &inter_planes,
                                 const csdo_params *vehicle = nullptr, int device = 0) {
  csdo_detail::Packed p;
  csdo_detail::pack_guess(x0_bar, p);
  const csdo_params P = detail::plane_params(vehicle, nullptr, nullptr);
  std::vector<int32_t> flat((size_t)3 * neighbor_pairs.size());
  for (size_t q = 0; q < neighbor_pairs.size(); ++q)
    for (int e = 0; e < 3; ++e) flat[3 * q + e] = neighbor_pairs[q][e];
  p.plane_t.assign(2 * neighbor_pairs.size(), 0);
  p.plane_abc.assign((size_t)24 * neighbor_pairs.size(), 0.0);
  {
    csdo_detail::HandleLease lease(P, device);
    csdo_batch b = p.view();
    csdo_detail::check(csdo_planes_from_pairs(lease.get(), &b, (int64_t)neighbor_pairs.size(), flat.data(),
                                              p.plane_ptr.data(), p.plane_t.data(), p.plane_abc.data()),
                       lease.get(), "csdo_planes_from_pairs");
  }
  detail::unpack_planes(p, inter_planes);
}

// both in one call (the pair list is only an intermediate of the reference, csdo.cc:119-129).
// Returns initial_inter_legal.
inline bool buildInterPlanes(const std::vector<std::vector<OptimizeResult>> &x0_bar,
                             std::vector<std::vector<InterPlane>> &inter_planes, const csdo_params *params = nullptr,
                             int device = 0) {
  csdo_detail::Packed p;
  const bool legal = detail::build(x0_bar, detail::plane_params(params, nullptr, nullptr), device, p, nullptr);
  detail::unpack_planes(p, inter_planes);
  return legal;
}

}  // namespace csdo_b200

namespace libMultiRobotPlanning {
using csdo_b200::buildInterPlanes;
#if !defined(CSDO_WITH_REFERENCE_TYPES)
using csdo_b200::calcEqualInterPlanes;
using csdo_b200::findNeighborPairsByTrustRegion;
#endif
}  // namespace libMultiRobotPlanning

class SolverDSQP {
 public:
  using OptimizeResult = libMultiRobotPlanning::OptimizeResult;
  using InterPlane = libMultiRobotPlanning::InterPlane;
  using QpParm = libMultiRobotPlanning::QpParm;
  using Corridor = libMultiRobotPlanning::Corridor;

  // Same argument list as sqp/dsqp_solver.h:26-34; `vehicle` (optional) replaces the reference's
  // global Constants (f2x, r2x, rv, WB, steer_max ...): nullptr = the shipped config.yaml.
  template <class ObstacleContainer>
  SolverDSQP(std::vector<std::vector<OptimizeResult>> &solutions,
             const std::vector<std::vector<OptimizeResult>> &x0_bar,
             const std::vector<std::vector<InterPlane>> &inter_planes, double dimx, double dimy,
             const ObstacleContainer &obstacles, const QpParm &param, int logger_level = 2,
             const csdo_params *vehicle = nullptr, int device = 0) {
    (void)logger_level;
    csdo_params P;
    if (vehicle) P = *vehicle; else csdo_default_params(&P);
    P.r_trust = param.r_trust; P.max_omega = param.max_omega; P.max_v = param.max_v;
    P.max_iter = (int)param.max_iter; P.delta_solution_threshold = param.delta_solution_threshold;
    P.osqp_max_iter = param.osqp_max_iter; P.fixed_corridor = param.fixed_corridor ? 1 : 0;
    if (param.dt > 0) P.dt = param.dt;
    csdo_detail::Packed p;
    csdo_detail::pack_guess(x0_bar, p);
    const int Na = p.inst_agent_ptr[1], Nt = p.inst_nt[0];
    p.dims[0] = dimx; p.dims[1] = dimy;
    for (const auto &o : obstacles) { p.obs.push_back(o.x); p.obs.push_back(o.y); p.obs.push_back(o.r); }
    p.obs_ptr[1] = (int32_t)(p.obs.size() / 3);
    for (int a = 0; a < Na; ++a) {
      p.plane_ptr[a + 1] = p.plane_ptr[a] + (int32_t)inter_planes[a].size();
      for (const InterPlane &q : inter_planes[a]) {
        p.plane_t.push_back(q.t);
        const double v[12] = {q.a_f2f, q.b_f2f, q.c_f2f, q.a_f2r, q.b_f2r, q.c_f2r,
                              q.a_r2f, q.b_r2f, q.c_r2f, q.a_r2r, q.b_r2r, q.c_r2r};
        p.plane_abc.insert(p.plane_abc.end(), v, v + 12);
      }
    }
    std::vector<double> traj((size_t)6 * Na * Nt), corr((size_t)8 * Na * Nt), obj(Na);
    std::vector<int32_t> status(Na), sqp(Na), nqp(Na), admm(Na), nfac(Na);
    int32_t inst_status = 0, inst_legal = 1;
    csdo_result r{traj.data(), corr.data(), status.data(), sqp.data(), nqp.data(), admm.data(), nfac.data(),
                  obj.data(), &inst_status, &inst_legal};
    {
      csdo_detail::HandleLease lease(P, device);
      const auto t0 = std::chrono::steady_clock::now();
      csdo_batch b = p.view();
      const int rc = csdo_refine(lease.get(), &b, &r);
      max_individual_opt_runtime = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      csdo_detail::check(rc, lease.get(), "csdo_refine");
    }
    solutions.assign(Na, {});
    corridors.assign(Na, {});
    num_iterations.assign(sqp.begin(), sqp.end());
    for (int a = 0; a < Na; ++a) {
      const double *g = traj.data() + (size_t)6 * a * Nt, *c = corr.data() + (size_t)8 * a * Nt;
      solutions[a].resize(Nt);
      corridors[a].resize(Nt);
      for (int t = 0; t < Nt; ++t) {
        OptimizeResult &o = solutions[a][t];
        o.x = g[t]; o.y = g[Nt + t]; o.yaw = g[2 * Nt + t]; o.steer = g[3 * Nt + t];
        o.v = g[4 * Nt + t]; o.d_steer = g[5 * Nt + t];
        corridors[a][t] = Corridor{c[t], c[Nt + t], c[2 * Nt + t], c[3 * Nt + t],
                                   c[4 * Nt + t], c[5 * Nt + t], c[6 * Nt + t], c[7 * Nt + t]};
      }
    }
    solve_status = inst_status;
    initial_static_legal = inst_legal != 0;
    agent_status.assign(status.begin(), status.end());
    admm_iterations.assign(admm.begin(), admm.end());
  }

  int getSolverStatus() { return solve_status; }
  // reference: max per-agent time + bookkeeping ("ideal parallel processing time"); here the agents
  // really run concurrently, so this is the wall time of the whole refine
  double getMaxOfRuntimes() { return max_individual_opt_runtime; }
  bool get_initial_static_legal() { return initial_static_legal; }

  std::vector<int> num_iterations;              // public in the reference (dsqp_solver.h:46)
  std::vector<std::vector<Corridor>> corridors; // public in the reference (dsqp_solver.h:47)
  std::vector<int> agent_status;                // extra: OSQP code of each agent's last QP
  std::vector<int> admm_iterations;             // extra: ADMM iterations per agent

 private:
  int solve_status = 0;
  double max_individual_opt_runtime = -1;
  bool initial_static_legal = true;
};
