// TEST INFRASTRUCTURE.  C entry points over the REFERENCE'S OWN corridor / inter-agent code, compiled
// unmodified from /root/reference/sqp/corridor.cc and sqp/inter_agent_cons.cc (oracle/Makefile target
// `ref`).  Used to pin the restatement (oracle/dsqp_restate.c) and to mint tests/golden fixtures; never
// loaded by the product path.  Nothing of the reference is copied: this file only calls its functions.
#include <array>
#include <cstring>
#include <string>
#include <unordered_set>
#include <vector>

#include "common/motion_planning.h"
#include "hybrid_a_star/planresult.h"
#include "sqp/common.h"
#include "sqp/corridor.h"
#include "sqp/inter_agent_cons.h"

using namespace libMultiRobotPlanning;

namespace {
std::unordered_set<Location> make_set(const double *obs, int n) {
  std::unordered_set<Location> s;
  for (int i = 0; i < n; ++i) s.insert(Location(obs[3 * i], obs[3 * i + 1], obs[3 * i + 2]));  // Instance.cc:41-47 order
  return s;
}
std::vector<std::vector<OptimizeResult>> make_guess(const double *g, int Na, int Nt) {
  std::vector<std::vector<OptimizeResult>> v(Na, std::vector<OptimizeResult>(Nt));
  for (int a = 0; a < Na; ++a)
    for (int t = 0; t < Nt; ++t) {
      const double *p = g + (size_t)a * 6 * Nt;
      OptimizeResult r{};
      r.x = p[t]; r.y = p[Nt + t]; r.yaw = p[2 * Nt + t]; r.steer = p[3 * Nt + t]; r.v = p[4 * Nt + t];
      r.d_steer = p[5 * Nt + t]; r.a = 0;
      v[a][t] = r;
    }
  return v;
}
}  // namespace

extern "C" {

// iteration order of std::unordered_set<Location> built by inserting obs[0..n) in order: out[k] = index
// of the k-th obstacle visited; returns the number of distinct obstacles
int ref_obstacle_order(const double *obs, int n, int *out) {
  const auto s = make_set(obs, n);
  int k = 0;
  for (const auto &o : s) {
    int idx = -1;
    for (int i = 0; i < n; ++i)
      if (obs[3 * i] == o.x && obs[3 * i + 1] == o.y && obs[3 * i + 2] == o.r) { idx = i; break; }
    out[k++] = idx;
  }
  return k;
}

// generateBox (corridor.cc:124-159): box = x_min,y_min,x_max,y_max; status = success, intial_status
void ref_generate_box(double dimx, double dimy, double x, double y, const double *obs, int n, double *box,
                      int *status) {
  const auto s = make_set(obs, n);
  Box b(0, 0, 0, 0);
  const BoxStatus st = generateBox(dimx, dimy, x, y, s, b);
  box[0] = b.x_min; box[1] = b.y_min; box[2] = b.x_max; box[3] = b.y_max;
  status[0] = st.success ? 1 : 0; status[1] = st.intial_status;
}

// calcCorridors (corridor.cc:164-248).  guess: Na x 6 planes of Nt; corr: Na x 8 planes of Nt in Corridor
// member order; returns initial_success
int ref_calc_corridors(const double *guess, int Na, int Nt, double dimx, double dimy, const double *obs, int n,
                       double *corr) {
  const auto s = make_set(obs, n);
  const auto g = make_guess(guess, Na, Nt);
  std::vector<std::vector<Corridor>> c;
  double tmax = 0;
  const bool ok = calcCorridors(g, s, dimx, dimy, c, Na, Nt, tmax, 0);
  for (int a = 0; a < Na; ++a)
    for (int t = 0; t < Nt; ++t) {
      double *p = corr + (size_t)a * 8 * Nt;
      const Corridor &k = c[a][t];
      p[t] = k.xf_min; p[Nt + t] = k.xf_max; p[2 * Nt + t] = k.yf_min; p[3 * Nt + t] = k.yf_max;
      p[4 * Nt + t] = k.xr_min; p[5 * Nt + t] = k.xr_max; p[6 * Nt + t] = k.yr_min; p[7 * Nt + t] = k.yr_max;
    }
  return ok ? 1 : 0;
}

// findNeighborPairsByTrustRegion + calcEqualInterPlanes (inter_agent_cons.cc:12-140).  Call with
// plane_t == NULL to get the per-agent counts (plane_cnt[Na]); then with plane_ptr (exclusive scan) to
// receive plane_t / plane_abc (12 doubles per plane, InterPlane member order).  *n_pairs: pair count.
int ref_instance_planes(const double *guess, int Na, int Nt, double r_trust, int *plane_cnt, int *plane_t,
                        double *plane_abc, const int *plane_ptr, int *n_pairs) {
  const auto g = make_guess(guess, Na, Nt);
  std::vector<std::array<int, 3>> pairs;
  const bool legal = findNeighborPairsByTrustRegion(g, r_trust, Constants::rv, pairs);
  std::vector<std::vector<InterPlane>> planes;
  calcEqualInterPlanes(g, pairs, planes);
  if (n_pairs) *n_pairs = (int)pairs.size();
  for (int a = 0; a < Na; ++a) {
    plane_cnt[a] = (int)planes[a].size();
    if (!plane_t) continue;
    for (size_t k = 0; k < planes[a].size(); ++k) {
      const InterPlane &q = planes[a][k];
      plane_t[plane_ptr[a] + k] = q.t;
      double *o = plane_abc + 12 * ((size_t)plane_ptr[a] + k);
      o[0] = q.a_f2f; o[1] = q.b_f2f; o[2] = q.c_f2f; o[3] = q.a_f2r; o[4] = q.b_f2r; o[5] = q.c_f2r;
      o[6] = q.a_r2f; o[7] = q.b_r2f; o[8] = q.c_r2f; o[9] = q.a_r2r; o[10] = q.b_r2r; o[11] = q.c_r2r;
    }
  }
  return legal ? 1 : 0;
}

// InterpolateInitalGuess (inter_agent_cons.cc:143-157) for Na coarse paths.  states: concatenated
// [sum n_states][3], actions: concatenated [sum (n_states-1)], n_states[Na]; goals [Na][3] or NULL (then
// the last coarse state is its own goal).  out: Na x 6 planes of nt_cap; returns the horizon Nt (<= nt_cap)
// or -1.
int ref_interpolate_guess(int Na, const int *n_states, const double *states, const int *actions,
                          const double *goals, int num_interpolation, double dt, int nt_cap, double *out) {
  std::vector<PlanResult<State, Action, double>> sol(Na);
  std::vector<State> gl;
  size_t so = 0, ao = 0;
  for (int a = 0; a < Na; ++a) {
    for (int i = 0; i < n_states[a]; ++i, ++so)
      sol[a].states.emplace_back(State(states[3 * so], states[3 * so + 1], states[3 * so + 2], i), (double)i);
    for (int i = 0; i + 1 < n_states[a]; ++i, ++ao) sol[a].actions.emplace_back(actions[ao], 1.0);
    const auto &last = sol[a].states.back().first;
    if (goals) gl.emplace_back(goals[3 * a], goals[3 * a + 1], goals[3 * a + 2]);
    else gl.emplace_back(last.x, last.y, last.yaw, last.time);
  }
  QpParm qp{};
  qp.num_interpolation = num_interpolation;
  qp.dt = dt;
  std::vector<std::vector<OptimizeResult>> x0;
  InterpolateInitalGuess(sol, x0, gl, qp);
  const int Nt = (int)x0[0].size();
  if (Nt > nt_cap) return -1;
  for (int a = 0; a < Na; ++a)
    for (int t = 0; t < Nt; ++t) {
      double *p = out + (size_t)a * 6 * nt_cap;
      const OptimizeResult &r = x0[a][t];
      p[t] = r.x; p[nt_cap + t] = r.y; p[2 * nt_cap + t] = r.yaw; p[3 * nt_cap + t] = r.steer;
      p[4 * nt_cap + t] = r.v; p[5 * nt_cap + t] = r.d_steer;
    }
  return Nt;
}

// dumpSolutions (inter_agent_cons.cc:413-455).  sol: Na x 6 planes of Nt; stat: the 10 SolutionStatistics
// fields in declaration order (sqp/common.h:25-36), the last two as ints.
void ref_dump_solutions(const char *file, const double *sol, int Na, int Nt, const double *stat) {
  const auto s = make_guess(sol, Na, Nt);
  SolutionStatistics st;
  st.cost = stat[0]; st.makespan = stat[1]; st.flowtime = stat[2]; st.runtime = stat[3]; st.rt_search = stat[4];
  st.rt_preprocess = stat[5]; st.rt_optimization = stat[6]; st.rt_max_optimization = stat[7];
  st.search_status = (int)stat[8]; st.solver_status = (int)stat[9];
  dumpSolutions(std::string(file), s, st);
}

}  // extern "C"
