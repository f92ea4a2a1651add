// Interpolated initial guess x0_bar of the DSQP refine stage, header-only C++.
//
// Restates the reference's InterpolateInitalGuess chain (sqp/inter_agent_cons.cc:143-411) on the host,
// in double precision, operation for operation: every coarse (state, action) step of the priority-based
// plan is split into num_interpolation + 1 sub-arcs whose radius is re-fitted to the chord (:214-241),
// yaw accumulates un-normalised (:265-267), the goal state is forced exact (:149-151); steer follows the
// action (:348-370), v is the projected displacement / dt and w the steer difference / dt (:389-398);
// agents that arrive early hold their pose with v = w = steer = 0 (:341-345, :368-370, :400-403).
// The PBS path types of the reference (pbs/) are outside this library: a coarse path is given as plain
// states + actions (0 fwd-straight, 1 fwd-right, 2 fwd-left, 3 rev-straight, 4, 5 rev turns, 6 wait;
// common/motion_planning.cc:47-51).
#pragma once

#include <cmath>
#include <vector>

#include "csdo/dsqp_solver.h"

namespace libMultiRobotPlanning {

struct CoarseState {
  double x, y, yaw;
};
struct CoarsePath {
  std::vector<CoarseState> states;  // n_actions + 1
  std::vector<int> actions;
};

namespace initial_guess_detail {
// motion_planning.h:70-75 (float return type)
inline double normalizeAngleAbsInPi(double x) {
  const double pi = 3.14159265358979323846;
  x = std::fmod(x + pi, 2 * pi);
  if (x < 0) x += 2 * pi;
  return (double)(float)(x - pi);
}

// interpolateXYYaw + action_sample (inter_agent_cons.cc:194-311)
inline void interpolatePath(const std::vector<CoarseState> &states, const std::vector<int> &actions, int n,
                            double r_const, std::vector<CoarseState> &fine, std::vector<int> &acts) {
  fine.clear(); acts.clear();
  fine.push_back(states[0]);
  CoarseState s = states[0];
  for (size_t i = 0; i < actions.size(); ++i) {
    const int action = actions[i];
    const CoarseState s0 = s, s1 = states[i + 1];
    for (int k = 0; k < n + 1; ++k) acts.push_back(action);
    if (action == 6) {
      for (int k = 0; k < n + 1; ++k) fine.push_back(s0);
    } else {
      double r = r_const, deltat;
      const double ddx = s1.x - s0.x, ddy = s1.y - s0.y;
      if (action == 0 || action == 3) {
        deltat = std::sqrt(ddx * ddx + ddy * ddy) / r_const;
      } else {
        deltat = normalizeAngleAbsInPi(s1.yaw - s0.yaw);
        const double d = std::sqrt(ddx * ddx + ddy * ddy);
        r = d / (2.0 * std::sin(std::fabs(deltat) / 2.0));
      }
      const double da = std::fabs(deltat) / (double)(n + 1);
      const double sx = r * std::sin(da), cy = r * (1 - std::cos(da));
      const double dxs[6] = {r * da, sx, sx, -r * da, -sx, -sx};
      const double dys[6] = {0.0, -cy, cy, 0.0, -cy, cy};
      const double dyaws[6] = {0.0, -da, da, 0.0, da, -da};
      const double dx = dxs[action], dy = dys[action], dyaw = dyaws[action];
      CoarseState c = s0;
      for (int k = 0; k < n; ++k) {
        const double xs = c.x + dx * std::cos(c.yaw) - dy * std::sin(c.yaw);
        const double ys = c.y + dx * std::sin(c.yaw) + dy * std::cos(c.yaw);
        c = CoarseState{xs, ys, c.yaw + dyaw};
        fine.push_back(c);
      }
      fine.push_back(CoarseState{s1.x, s1.y, ((action == 0 || action == 3) ? 0.0 : deltat) + s0.yaw});
    }
    s = fine.back();
  }
}
}  // namespace initial_guess_detail

// InterpolateInitalGuess (inter_agent_cons.cc:143-157): all agents padded to the longest horizon.
// goals may be null (then the last coarse state is kept).  dt, LF, LB as in csdo_params.
inline void InterpolateInitalGuess(const std::vector<CoarsePath> &solution,
                                   std::vector<std::vector<OptimizeResult>> &x0_bar,
                                   const std::vector<CoarseState> *goals, double dt, double LF, double LB,
                                   int num_interpolation = 2, double r_const = 3.0) {
  using namespace initial_guess_detail;
  const size_t na = solution.size();
  std::vector<std::vector<CoarseState>> fine(na);
  std::vector<std::vector<int>> acts(na);
  size_t nt = 0;
  for (size_t a = 0; a < na; ++a) {
    std::vector<CoarseState> st = solution[a].states;
    if (goals) st.back() = (*goals)[a];  // :149-151
    interpolatePath(st, solution[a].actions, num_interpolation, r_const, fine[a], acts[a]);
    nt = fine[a].size() > nt ? fine[a].size() : nt;
  }
  const double phi = (double)std::atan((float)(((float)LF - (float)LB) / (float)r_const));  // std::atan(float), :352-353
  x0_bar.assign(na, std::vector<OptimizeResult>(nt));
  for (size_t a = 0; a < na; ++a) {
    const size_t ns = fine[a].size();
    std::vector<OptimizeResult> &g = x0_bar[a];
    for (size_t t = 0; t < nt; ++t) {
      const CoarseState &s = fine[a][t < ns ? t : ns - 1];
      g[t] = OptimizeResult{};
      g[t].x = s.x; g[t].y = s.y; g[t].yaw = s.yaw;
    }
    for (size_t i = 1; i < ns; ++i) {
      const int act = acts[a][i - 1];
      g[i].steer = (act == 0 || act == 3 || act == 6) ? 0.0 : ((act == 1 || act == 4) ? -phi : phi);
    }
    for (size_t t = 0; t + 1 < ns; ++t) {
      g[t].v = ((g[t + 1].x - g[t].x) / dt) * std::cos(g[t].yaw) + ((g[t + 1].y - g[t].y) / dt) * std::sin(g[t].yaw);
      g[t].d_steer = (g[t + 1].steer - g[t].steer) / dt;
    }
  }
}

}  // namespace libMultiRobotPlanning
