// CPU-only caller of include/csdo/solution_io.h: reads "na nt solver_status search_status rt_preprocess" and
// na*nt rows "x y yaw steer v w" from stdin, writes the solution file named by argv[1].
#include <cstdio>
#include <vector>

#include "csdo/solution_io.h"

using namespace libMultiRobotPlanning;

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  int na, nt;
  SolutionStatistics st;
  if (std::scanf("%d %d %d %d %lf", &na, &nt, &st.solver_status, &st.search_status, &st.rt_preprocess) != 5) return 1;
  std::vector<std::vector<OptimizeResult>> sol(na, std::vector<OptimizeResult>(nt));
  for (auto &ag : sol)
    for (auto &s : ag)
      if (std::scanf("%lf %lf %lf %lf %lf %lf", &s.x, &s.y, &s.yaw, &s.steer, &s.v, &s.d_steer) != 6) return 1;
  dumpSolutions(argv[1], sol, st);
  return 0;
}
