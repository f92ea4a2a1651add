// CPU-only caller of include/csdo/initial_guess.h: reads coarse paths from stdin, prints x0_bar.
//   input : na  dt LF LB  then per agent: n_actions, (n_actions+1) x "x y yaw", n_actions x action, goal "x y yaw"
//   output: nt, then per agent and step "x y yaw steer v w" with 17 significant digits
#include <cstdio>
#include <vector>

#include "csdo/initial_guess.h"

using namespace libMultiRobotPlanning;

int main() {
  int na;
  double dt, LF, LB;
  if (std::scanf("%d %lf %lf %lf", &na, &dt, &LF, &LB) != 4) return 1;
  std::vector<CoarsePath> sol(na);
  std::vector<CoarseState> goals(na);
  for (int a = 0; a < na; ++a) {
    int n;
    if (std::scanf("%d", &n) != 1) return 1;
    sol[a].states.resize(n + 1);
    sol[a].actions.resize(n);
    for (auto &s : sol[a].states)
      if (std::scanf("%lf %lf %lf", &s.x, &s.y, &s.yaw) != 3) return 1;
    for (auto &k : sol[a].actions)
      if (std::scanf("%d", &k) != 1) return 1;
    if (std::scanf("%lf %lf %lf", &goals[a].x, &goals[a].y, &goals[a].yaw) != 3) return 1;
  }
  std::vector<std::vector<OptimizeResult>> x0;
  InterpolateInitalGuess(sol, x0, &goals, dt, LF, LB);
  std::printf("%zu\n", x0.empty() ? (size_t)0 : x0[0].size());
  for (auto &ag : x0)
    for (auto &s : ag) std::printf("%.17g %.17g %.17g %.17g %.17g %.17g\n", s.x, s.y, s.yaw, s.steer, s.v, s.d_steer);
  return 0;
}
