// CPU-only caller of include/csdo/solution_io.h: reads "na nt solver_status search_status rt_preprocess" and
// na*nt rows "x y yaw steer v w" from stdin, writes the solution file named by argv[1].  With a second
// argument "full" the header is "na nt" followed by all 10 SolutionStatistics values in declaration order.
#include <cstdio>
#include <string>
#include <vector>

#include "csdo/solution_io.h"

using namespace libMultiRobotPlanning;

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  int na, nt;
  SolutionStatistics st;
  if (argc > 2 && std::string(argv[2]) == "full") {
    double ss, so;
    if (std::scanf("%d %d %lf %lf %lf %lf %lf %lf %lf %lf %lf %lf", &na, &nt, &st.cost, &st.makespan, &st.flowtime, &st.runtime,
                   &st.rt_search, &st.rt_preprocess, &st.rt_optimization, &st.rt_max_optimization, &ss, &so) != 12) return 1;
    st.search_status = (int)ss; st.solver_status = (int)so;
  } else if (std::scanf("%d %d %d %d %lf", &na, &nt, &st.solver_status, &st.search_status, &st.rt_preprocess) != 5) return 1;
  std::vector<std::vector<OptimizeResult>> sol(na, std::vector<OptimizeResult>(nt));
  for (auto &ag : sol)
    for (auto &s : ag)
      if (std::scanf("%lf %lf %lf %lf %lf %lf", &s.x, &s.y, &s.yaw, &s.steer, &s.v, &s.d_steer) != 6) return 1;
  dumpSolutions(argv[1], sol, st);
  return 0;
}
