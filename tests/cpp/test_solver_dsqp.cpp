// C++ caller of the drop-in interface, shaped like the reference's csdo.cc:113-167:
// x0_bar -> planes (findNeighborPairsByTrustRegion + calcEqualInterPlanes) -> SolverDSQP -> print.
// usage: test_solver_dsqp <guess.txt>   (Na Nt dimx dimy No, then obstacles, then Na*Nt rows x y yaw steer v w)
// Output: one line per agent "a status sqp_iters x_t... " consumed by tests/test_gpu_parity.py.
#include <cstdio>
#include <fstream>
#include <iostream>
#include <unordered_set>

#include "csdo/dsqp_solver.h"

using namespace libMultiRobotPlanning;

int main(int argc, char **argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: %s guess.txt\n", argv[0]); return 2; }
  std::ifstream in(argv[1]);
  int Na, Nt, No;
  double dimx, dimy;
  in >> Na >> Nt >> dimx >> dimy >> No;
  // the reference's own container type (sqp/dsqp_solver.h:31); its iteration order is printed so that the
  // checker can hand the same order to the C ABI
  std::unordered_set<Location> obstacles;
  std::vector<Location> listed;
  for (int o = 0; o < No; ++o) { double x, y, r; in >> x >> y >> r; obstacles.insert(Location(x, y, r)); listed.emplace_back(x, y, r); }
  if (obstacles.size() != listed.size()) { std::fprintf(stderr, "duplicate obstacles\n"); return 2; }
  std::vector<std::vector<OptimizeResult>> x0_bar(Na, std::vector<OptimizeResult>(Nt));
  for (int a = 0; a < Na; ++a)
    for (int t = 0; t < Nt; ++t) {
      OptimizeResult &r = x0_bar[a][t];
      in >> r.x >> r.y >> r.yaw >> r.steer >> r.v >> r.d_steer;
    }
  try {
    QpParm param;  // readQpSolverConfig defaults of the shipped config.yaml; dt = 0 -> library default
    std::vector<std::vector<InterPlane>> inter_planes;
    // the reference's call sequence (csdo.cc:119-129) ...
    std::vector<std::array<int, 3>> neighbor_pairs;
    const bool initial_inter_legal = findNeighborPairsByTrustRegion(x0_bar, param.r_trust, 1.25, neighbor_pairs);
    calcEqualInterPlanes(x0_bar, neighbor_pairs, inter_planes);
    // ... gives the same planes as the fused call
    std::vector<std::vector<InterPlane>> fused;
    const bool legal2 = buildInterPlanes(x0_bar, fused);
    bool same = legal2 == initial_inter_legal && fused.size() == inter_planes.size();
    size_t n_planes = 0;
    for (size_t a = 0; same && a < fused.size(); ++a) {
      same = fused[a].size() == inter_planes[a].size();
      n_planes += fused[a].size();
      for (size_t k = 0; same && k < fused[a].size(); ++k)
        same = std::memcmp(&fused[a][k], &inter_planes[a][k], sizeof(InterPlane)) == 0;
    }
    if (!same || n_planes != 2 * neighbor_pairs.size()) { std::fprintf(stderr, "pair-list planes differ from the fused build\n"); return 3; }
    for (size_t q = 1; q < neighbor_pairs.size(); ++q)
      if (!(neighbor_pairs[q - 1] < neighbor_pairs[q])) { std::fprintf(stderr, "pairs not in (t,i,j) order\n"); return 3; }
    std::vector<std::vector<OptimizeResult>> optimize_res;
    SolverDSQP solver(optimize_res, x0_bar, inter_planes, dimx, dimy, obstacles, param, 0);
    std::printf("obstacle_order");
    for (const Location &o : obstacles)
      for (size_t q = 0; q < listed.size(); ++q)
        if (listed[q] == o) std::printf(" %zu", q);
    std::printf("\n");
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
