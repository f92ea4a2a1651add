// Solution files of the refine stage, header-only C++: the YAML layout of the
// reference's dumpSolutions (sqp/inter_agent_cons.cc:413-455) and its statistics record
// (sqp/common.h:25-36), so that a caller of the drop-in SolverDSQP can write files that the reference's
// scripts/analysis_result.py and scripts/visualize.py read unchanged.
#pragma once

#include <cmath>
#include <fstream>
#include <iomanip>
#include <string>
#include <vector>

#include "csdo/dsqp_solver.h"

namespace libMultiRobotPlanning {

struct SolutionStatistics {  // sqp/common.h:25-36
  double cost = -1;
  double makespan = -1;
  double flowtime = -1;
  double runtime = -1;
  double rt_search = -1;
  double rt_preprocess = -1;
  double rt_optimization = -1;
  double rt_max_optimization = -1;
  int search_status = 2;  // 0: failed; 1: minor collision; 2: success
  int solver_status = 0;  // SolverDSQP::getSolverStatus()
};

// Fixed 3 decimals; steer and omega are written as value * 180 / 3.14; the last step carries no v / omega.
inline void dumpSolutions(const std::string &file_name, const std::vector<std::vector<OptimizeResult>> &solutions,
                          const SolutionStatistics &stat) {
  std::ofstream out(file_name);
  const size_t Na = solutions.size();
  const size_t Nt = Na ? solutions[0].size() : 0;
  out << std::fixed << std::setprecision(3);
  out << "statistics:\n";
  out << "  cost: " << stat.cost << "\n";
  out << "  makespan: " << stat.makespan << "\n";
  out << "  flowtime: " << stat.flowtime << "\n";
  out << "  runtime: " << stat.runtime << "\n";
  out << "  runtime_search: " << stat.rt_search << "\n";
  out << "  runtime_preprocess: " << stat.rt_preprocess << "\n";
  out << "  runtime_optimization: " << stat.rt_optimization << "\n";
  out << "  runtime_decentralized_optimization: " << stat.rt_max_optimization << "\n";
  out << "  search_status: " << stat.search_status << "\n";
  out << "  solver_status: " << stat.solver_status << "\n";
  out << "schedule:\n";
  for (size_t a = 0; a < Na; ++a) {
    out << "  agent" << a << ":\n";
    for (size_t t = 0; t < solutions[a].size(); ++t) {
      const OptimizeResult &s = solutions[a][t];
      out << "    - x: " << s.x << "\n"
          << "      y: " << s.y << "\n"
          << "      yaw: " << s.yaw << "\n"
          << "      steer: " << s.steer * 180 / 3.14 << "\n"
          << "      t: " << t << "\n";
      if (t == Nt - 1) continue;
      out << "      v: " << s.v << "\n"
          << "      omega: " << s.d_steer * 180 / 3.14 << "\n";
    }
  }
}

// Corridor file of `./csdo --dump_corridor` (dumpCorridors, sqp/utils.cc:62-89): per agent and step two rows
// "[disc centre x, y, x_min, x_max, y_min, y_max]" (front disc, rear disc) of the INITIAL GUESS x0_bar, in the
// stream's default 6-significant-digit format; the centres go through State's float members
// (common/motion_planning.h:115-118, 201-206).  f2x / r2x: the float Constants (csdo_params::f2x, r2x).
inline void dumpCorridors(const std::string &file_name, const std::vector<std::vector<Corridor>> &corridors,
                          const std::vector<std::vector<OptimizeResult>> &x0_bar, double f2x = 1.25, double r2x = -0.25) {
  std::ofstream out(file_name);
  const size_t Na = corridors.size();
  const size_t Nt = Na ? corridors[0].size() : 0;
  for (size_t a = 0; a < Na; ++a) {
    out << "agent" << a << ":" << std::endl;
    for (size_t t = 0; t < Nt; ++t) {
      const OptimizeResult &s = x0_bar[a][t];
      const Corridor &c = corridors[a][t];
      const double xf = (float)(s.x + (float)f2x * std::cos(s.yaw)), yf = (float)(s.y + (float)f2x * std::sin(s.yaw));
      const double xr = (float)(s.x + (float)r2x * std::cos(s.yaw)), yr = (float)(s.y + (float)r2x * std::sin(s.yaw));
      out << "  - [" << xf << ", " << yf << ", " << c.xf_min << ", " << c.xf_max << ", " << c.yf_min << ", " << c.yf_max << "]\n";
      out << "  - [" << xr << ", " << yr << ", " << c.xr_min << ", " << c.xr_max << ", " << c.yr_min << ", " << c.yr_max << "]\n";
    }
  }
}

}  // namespace libMultiRobotPlanning
