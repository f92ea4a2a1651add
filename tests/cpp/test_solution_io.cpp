// CPU-only caller of include/csdo/solution_io.h: reads "na nt solver_status search_status rt_preprocess" and
// na*nt rows "x y yaw steer v w" from stdin, writes the solution file named by argv[1].  With a second
// argument "full" the header is "na nt" followed by all 10 SolutionStatistics values in declaration order.
#include <cstdio>
#include <string>
#include <vector>

#include "csdo/solution_io.h"

using namespace libMultiRobotPlanning;

// "corridors" mode: "na nt f2x r2x", then na*nt rows "x y yaw" + 8 corridor values -> dumpCorridors
static int corridors_mode(const char *path) {
  int na, nt;
  double f2x, r2x;
  if (std::scanf("%d %d %lf %lf", &na, &nt, &f2x, &r2x) != 4) return 1;
  std::vector<std::vector<OptimizeResult>> x0(na, std::vector<OptimizeResult>(nt));
  std::vector<std::vector<Corridor>> co(na, std::vector<Corridor>(nt));
  for (int a = 0; a < na; ++a)
    for (int t = 0; t < nt; ++t) {
      OptimizeResult &s = x0[a][t];
      Corridor &c = co[a][t];
      if (std::scanf("%lf %lf %lf %lf %lf %lf %lf %lf %lf %lf %lf", &s.x, &s.y, &s.yaw, &c.xf_min, &c.xf_max, &c.yf_min,
                     &c.yf_max, &c.xr_min, &c.xr_max, &c.yr_min, &c.yr_max) != 11) return 1;
    }
  dumpCorridors(path, co, x0, f2x, r2x);
  return 0;
}

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  if (argc > 2 && std::string(argv[2]) == "corridors") return corridors_mode(argv[1]);
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
