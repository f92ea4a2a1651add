// C++ caller of the drop-in interface, shaped like the reference's csdo.cc:113-167:
// x0_bar -> planes (findNeighborPairsByTrustRegion + calcEqualInterPlanes) -> SolverDSQP -> print.
// usage: test_solver_dsqp <guess.txt>   (Na Nt dimx dimy No, then obstacles, then Na*Nt rows x y yaw steer v w)
// Output: one line per agent "a status sqp_iters x_t... " consumed by tests/test_gpu_parity.py.
#include <cstdio>
#include <fstream>
#include <iostream>
#include <unordered_set>

#include "csdo/dsqp_solver.h"

namespace std {
template <> struct hash<Location> {
  size_t operator()(const Location &s) const { return std::hash<double>()(s.x) * 31 + std::hash<double>()(s.y); }
};
}  // namespace std

using namespace libMultiRobotPlanning;

int main(int argc, char **argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: %s guess.txt\n", argv[0]); return 2; }
  std::ifstream in(argv[1]);
  int Na, Nt, No;
  double dimx, dimy;
  in >> Na >> Nt >> dimx >> dimy >> No;
  std::vector<Location> obstacles;  // any iterable of Location works; the reference uses unordered_set<Location>
  for (int o = 0; o < No; ++o) { double x, y, r; in >> x >> y >> r; obstacles.emplace_back(x, y, r); }
  std::vector<std::vector<OptimizeResult>> x0_bar(Na, std::vector<OptimizeResult>(Nt));
  for (int a = 0; a < Na; ++a)
    for (int t = 0; t < Nt; ++t) {
      OptimizeResult &r = x0_bar[a][t];
      in >> r.x >> r.y >> r.yaw >> r.steer >> r.v >> r.d_steer;
    }
  try {
    QpParm param;  // readQpSolverConfig defaults of the shipped config.yaml; dt = 0 -> library default
    std::vector<std::vector<InterPlane>> inter_planes;
    const bool initial_inter_legal = buildInterPlanes(x0_bar, inter_planes);
    std::vector<std::vector<OptimizeResult>> optimize_res;
    SolverDSQP solver(optimize_res, x0_bar, inter_planes, dimx, dimy, obstacles, param, 0);
    std::printf("solver_status %d static_legal %d inter_legal %d runtime %.6f\n", solver.getSolverStatus(),
                (int)solver.get_initial_static_legal(), (int)initial_inter_legal, solver.getMaxOfRuntimes());
    for (int a = 0; a < Na; ++a) {
      std::printf("agent %d %d %d %zu", a, solver.agent_status[a], solver.num_iterations[a], inter_planes[a].size());
      for (int t = 0; t < Nt; ++t) std::printf(" %.17g %.17g", optimize_res[a][t].x, solver.corridors[a][t].xf_min);
      std::printf("\n");
    }
  } catch (const std::exception &e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
