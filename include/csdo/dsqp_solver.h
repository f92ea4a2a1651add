// Drop-in C++ interface of the DSQP refine stage on top of the C ABI
// (include/csdo_dsqp.h).  Header-only; link with libcsdo_dsqp.so.
//
// It keeps the reference's own class and type names for this path so that the
// caller in csdo.cc:113-167 compiles unchanged against it:
//   OptimizeResult, QpParm            sqp/common.h:14-22, 39-52
//   Corridor                          sqp/corridor.h:8-11
//   InterPlane                        sqp/inter_agent_cons.h:47-63
//   Location                          common/motion_planning.h:80-97
//   findNeighborPairsByTrustRegion    sqp/inter_agent_cons.h:40-45
//   calcEqualInterPlanes              sqp/inter_agent_cons.h:70-73
//   SolverDSQP                        sqp/dsqp_solver.h:24-47 (constructor does the work)
// Differences that a maintainer has to know are listed in INTEGRATION.md: no
// Eigen/OSQP types (solveOSQP is gone), obstacles may be any iterable of
// Location (std::unordered_set<Location> included), and one extra optional
// constructor argument carries the vehicle constants that the reference keeps
// in the global `Constants` class.
#pragma once

#include <array>
#include <chrono>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../csdo_dsqp.h"

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
  bool operator==(const Location &o) const { return x == o.x && y == o.y && r == o.r; }
};

namespace csdo_detail {

using libMultiRobotPlanning::InterPlane;
using libMultiRobotPlanning::OptimizeResult;

inline void check(int rc, csdo_handle *h, const char *what) {
  if (rc != CSDO_OK) throw std::runtime_error(std::string(what) + ": " + (h ? csdo_last_error(h) : "no handle"));
}

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

namespace libMultiRobotPlanning {

// findNeighborPairsByTrustRegion + calcEqualInterPlanes in one call (the pair list itself is only an
// intermediate of the reference, csdo.cc:119-129).  Returns initial_inter_legal.
inline bool buildInterPlanes(const std::vector<std::vector<OptimizeResult>> &x0_bar,
                             std::vector<std::vector<InterPlane>> &inter_planes, const csdo_params *params = nullptr,
                             int device = 0) {
  csdo_detail::Packed p;
  csdo_detail::pack_guess(x0_bar, p);
  csdo_handle *h = nullptr;
  csdo_detail::check(csdo_create(params, device, &h), nullptr, "csdo_create");
  const int Na = p.inst_agent_ptr[1];
  int32_t legal = 1;
  csdo_batch b = p.view();
  try {
    csdo_detail::check(csdo_planes_count(h, &b, p.plane_ptr.data(), &legal), h, "csdo_planes_count");
    const int total = p.plane_ptr[Na];
    p.plane_t.assign(total, 0);
    p.plane_abc.assign((size_t)12 * total, 0.0);
    csdo_detail::check(csdo_planes_fill(h, &b, p.plane_ptr.data(), p.plane_t.data(), p.plane_abc.data()), h,
                       "csdo_planes_fill");
  } catch (...) { csdo_destroy(h); throw; }
  csdo_destroy(h);
  inter_planes.assign(Na, {});
  for (int a = 0; a < Na; ++a)
    for (int k = p.plane_ptr[a]; k < p.plane_ptr[a + 1]; ++k) {
      const double *q = p.plane_abc.data() + (size_t)12 * k;
      inter_planes[a].push_back(InterPlane{p.plane_t[k], q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7], q[8], q[9],
                                           q[10], q[11]});
    }
  return legal != 0;
}

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
    csdo_handle *h = nullptr;
    csdo_detail::check(csdo_create(&P, device, &h), nullptr, "csdo_create");
    const auto t0 = std::chrono::steady_clock::now();
    csdo_batch b = p.view();
    const int rc = csdo_refine(h, &b, &r);
    max_individual_opt_runtime = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (rc != CSDO_OK) { std::string e = csdo_last_error(h); csdo_destroy(h); throw std::runtime_error("csdo_refine: " + e); }
    csdo_destroy(h);
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
