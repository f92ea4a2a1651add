// "f1-lite": a deterministic prioritized planner over the reference planner's own seven motion primitives
// (common/motion_planning.cc:47-51, 96-108: forward / reverse straight, right, left with r = 3,
// deltat = 0.706, plus wait) from each benchmark scenario's start to its goal among the real obstacles.
// It stands in for the reference's PBS + spatiotemporal Hybrid A* front end (pbs/, hybrid_a_star/: host
// search, out of scope, needs OMPL) so that the REAL benchmark geometry can be pushed through
// InterpolateInitalGuess -> planes -> DSQP.  It is NOT a restatement of PBS: priorities are the agent
// order, there is no Reeds-Shepp shot, ties are broken deterministically.  Test / measurement
// infrastructure, host only.
//
// Space-time A* per agent: state (x, y, yaw, step); successors = the 6 primitives + wait; cost as in the
// reference's environment (turning x1.5, reversing x2, change of direction +2; config.yaml:8-13);
// validity: the reference planner's own rule (environment.h:350-392), see `valid` below.  The goal is reached within (1.2 m, 0.36 rad); the last state is then replaced by the
// exact goal, which is what InterpolateInitalGuess does anyway (inter_agent_cons.cc:149-151).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <queue>
#include <unordered_map>
#include <vector>

constexpr double kDt = 0.706;
namespace {
constexpr double kR = 3.0;
struct Pose { double x, y, yaw; };
struct Node { Pose s; int step, action, parent; double g; };
struct QItem { double f; int id; bool operator<(const QItem &o) const { return f > o.f || (f == o.f && id > o.id); } };

// primitive `a` over the arc angle kDt * scale (scale 1 = the planner's primitive; near the goal the search also
// tries half primitives: InterpolateInitalGuess re-fits every action's arc from its chord, so any length works)
Pose apply(const Pose &s, int a, double scale = 1.0) {
  if (a == 6) return s;
  const double kDt = ::kDt * scale;
  const double sx = kR * std::sin(kDt), cy = kR * (1 - std::cos(kDt));
  const double dxs[6] = {kR * kDt, sx, sx, -kR * kDt, -sx, -sx};
  const double dys[6] = {0.0, -cy, cy, 0.0, -cy, cy};
  const double dyaws[6] = {0.0, -kDt, kDt, 0.0, kDt, -kDt};
  const double c = std::cos(s.yaw), sn = std::sin(s.yaw);
  return Pose{s.x + dxs[a] * c - dys[a] * sn, s.y + dxs[a] * sn + dys[a] * c, s.yaw + dyaws[a]};
}
double wrap(double a) { a = std::fmod(a + M_PI, 2 * M_PI); if (a < 0) a += 2 * M_PI; return a - M_PI; }
}  // namespace

extern "C" {

// starts/goals [n_agents][3]; obstacles [n_obs][3].  Outputs: n_states[a] (0 = no path found), states
// [n_agents][max_states][3], actions [n_agents][max_states].  Returns the number of agents WITHOUT a path.
int plan_prioritized(double dimx, double dimy, int n_obs, const double *obs, int n_agents, const double *starts,
                     const double *goals, double f2x, double r2x, double rv, int max_states, int max_expansions,
                     int *n_states, double *states, int *actions) {
  std::vector<std::vector<Pose>> planned;  // higher-priority paths
  auto discs = [&](const Pose &s, double *d) {
    const double c = std::cos(s.yaw), sn = std::sin(s.yaw);
    d[0] = s.x + f2x * c; d[1] = s.y + f2x * sn; d[2] = s.x + r2x * c; d[3] = s.y + r2x * sn;
  };
  // the reference planner's own validity rule (hybrid_a_star/environment.h:350-392): disc centres inside
  // [rv, dim - rv], no obstacle centre inside the vehicle rectangle inflated by 1.2 r (State::obsCollision,
  // motion_planning.h:185-197), no rectangle overlap (State::agentCollision, SAT, :140-183) with a
  // higher-priority agent at steps t-1, t, t+1 (parked at its goal after its last step)
  const double LF = 2.0, LB = 1.0, W = 2.0;
  auto sat = [&](const Pose &a, const Pose &b) {
    const double d = (LF + LB) / 2 - LB, hl = (LF + LB) / 2, hw = W / 2;
    const double ca = std::cos(a.yaw), sa = std::sin(a.yaw), cb = std::cos(b.yaw), sb = std::sin(b.yaw);
    const double sx = (b.x + d * cb) - (a.x + d * ca), sy = (b.y + d * sb) - (a.y + d * sa);
    const double dx1 = ca * hl, dy1 = sa * hl, dx2 = sa * hw, dy2 = -ca * hw;
    const double dx3 = cb * hl, dy3 = sb * hl, dx4 = sb * hw, dy4 = -cb * hw;
    return std::fabs(sx * ca + sy * sa) <= std::fabs(dx3 * ca + dy3 * sa) + std::fabs(dx4 * ca + dy4 * sa) + hl &&
           std::fabs(sx * sa - sy * ca) <= std::fabs(dx3 * sa - dy3 * ca) + std::fabs(dx4 * sa - dy4 * ca) + hw &&
           std::fabs(sx * cb + sy * sb) <= std::fabs(dx1 * cb + dy1 * sb) + std::fabs(dx2 * cb + dy2 * sb) + hl &&
           std::fabs(sx * sb - sy * cb) <= std::fabs(dx1 * sb - dy1 * cb) + std::fabs(dx2 * sb - dy2 * cb) + hw;
  };
  auto valid = [&](const Pose &s, int step) {
    double d[4];
    discs(s, d);
    if (d[0] < rv || d[0] > dimx - rv || d[1] < rv || d[1] > dimy - rv || d[2] < rv || d[2] > dimx - rv || d[3] < rv ||
        d[3] > dimy - rv)
      return false;
    const double c = std::cos(s.yaw), sn = std::sin(s.yaw);
    for (int o = 0; o < n_obs; ++o) {
      const double ox = obs[3 * o] - s.x, oy = obs[3 * o + 1] - s.y, r = obs[3 * o + 2];
      const double rx = ox * c + oy * sn, ry = -ox * sn + oy * c;
      if (rx > -LB - r * 1.2 && rx < LF + r * 1.2 && ry > -W / 2.0 - r * 1.2 && ry < W / 2.0 + r * 1.2) return false;
    }
    for (const auto &p : planned)
      for (int dt = -1; dt <= 1; ++dt) {
        const int k = step + dt;
        if (k < 0) continue;
        if (sat(s, p[std::min<size_t>(k, p.size() - 1)])) return false;
      }
    return true;
  };
  int failed = 0;
  for (int a = 0; a < n_agents; ++a) {
    const Pose start{starts[3 * a], starts[3 * a + 1], starts[3 * a + 2]}, goal{goals[3 * a], goals[3 * a + 1], goals[3 * a + 2]};
    std::vector<Node> nodes;
    std::priority_queue<QItem> open;
    std::unordered_map<uint64_t, double> best;
    auto key = [&](const Pose &s, int step) {
      const int64_t ix = (int64_t)std::floor(s.x / 0.7), iy = (int64_t)std::floor(s.y / 0.7);
      const int64_t iw = (int64_t)std::floor((wrap(s.yaw) + M_PI) / (kDt / 2));
      return (uint64_t)((((ix * 4096 + iy) * 64 + iw) * 1024) + std::min(step, 1023));
    };
    auto heur = [&](const Pose &s) {
      return std::hypot(s.x - goal.x, s.y - goal.y) + 0.5 * std::fabs(wrap(s.yaw - goal.yaw));
    };
    nodes.push_back(Node{start, 0, -1, -1, 0.0});
    open.push(QItem{heur(start), 0});
    int found = -1, expansions = 0;
    // the goal pose must stay free of the higher-priority agents after arrival (they are parked or moving)
    while (!open.empty() && expansions < max_expansions) {
      const int id = open.top().id;
      open.pop();
      const Node n = nodes[id];
      ++expansions;
      if (std::hypot(n.s.x - goal.x, n.s.y - goal.y) < 1.2 && std::fabs(wrap(n.s.yaw - goal.yaw)) < 0.36 && n.step > 0) {
        bool stays_free = true;
        for (int k = 1; k <= 60 && stays_free; k += 3) stays_free = valid(goal, n.step + k);
        if (stays_free) { found = id; break; }
      }
      if (n.step + 1 >= max_states) continue;
      const bool near_goal = false;  // (half-length primitives near the goal were tried: more branching, fewer routed agents)
      for (int e = 0; e < (near_goal ? 13 : 7); ++e) {
        const int act = e < 7 ? e : e - 7;             // e >= 7: the six moving primitives at half length
        const double scale = e < 7 ? 1.0 : 0.5;
        const Pose s2 = apply(n.s, act, scale);
        if (!valid(s2, n.step + 1)) continue;
        double cost = kR * kDt * scale;
        if (act == 1 || act == 2 || act == 4 || act == 5) cost *= 1.5;
        if (act >= 3 && act < 6) cost *= 2.0;
        if (n.action >= 0 && n.action < 6 && act < 6 && ((n.action < 3) != (act < 3))) cost += 2.0;
        const double g2 = n.g + cost;
        const uint64_t k2 = key(s2, n.step + 1);
        auto it = best.find(k2);
        if (it != best.end() && it->second <= g2) continue;
        best[k2] = g2;
        nodes.push_back(Node{s2, n.step + 1, act, id, g2});
        open.push(QItem{g2 + 2.0 * heur(s2), (int)nodes.size() - 1});
      }
    }
    if (found < 0) {
      n_states[a] = 0;
      ++failed;
      planned.push_back(std::vector<Pose>{start});  // treated as parked at its start
      continue;
    }
    std::vector<Pose> path;
    std::vector<int> acts;
    for (int id = found; id >= 0; id = nodes[id].parent) { path.push_back(nodes[id].s); if (nodes[id].parent >= 0) acts.push_back(nodes[id].action); }
    std::reverse(path.begin(), path.end());
    std::reverse(acts.begin(), acts.end());
    path.back() = Pose{goal.x, goal.y, path.back().yaw + wrap(goal.yaw - path.back().yaw)};  // exact goal, continuous yaw
    n_states[a] = (int)path.size();
    for (size_t i = 0; i < path.size(); ++i) {
      double *o = states + ((size_t)a * max_states + i) * 3;
      o[0] = path[i].x; o[1] = path[i].y; o[2] = path[i].yaw;
      if (i < acts.size()) actions[(size_t)a * max_states + i] = acts[i];
    }
    planned.push_back(path);
  }
  return failed;
}

}  // extern "C"
