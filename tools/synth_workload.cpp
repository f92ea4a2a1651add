// Seeded synthetic coarse plans for bench.py / the full-size tests (measurement infrastructure, not
// part of the product path).  Same construction as csdotrajectoryplanning_b200/scenario.py::
// synthetic_instance -- every agent follows a random sequence of the planner's own motion primitives
// (common/motion_planning.cc:47-51,96-108: step r*deltat = 2.118 m, turn deltat = 0.706 rad at r = 3)
// that avoids the obstacles, the map border and the agents planned before it (a priority-style plan;
// the PBS + Hybrid-A* front end of the reference is out of scope) -- but in C++ with its own generator,
// so that BASELINE configs[4] (4096 instances x 100 agents) is produced in seconds.  The interpolated
// guess comes from include/csdo/initial_guess.h (InterpolateInitalGuess, inter_agent_cons.cc:143-411).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "csdo/initial_guess.h"

namespace {

struct Rng {  // xoshiro256** seeded through splitmix64
  uint64_t s[4];
  static uint64_t splitmix(uint64_t &x) {
    uint64_t z = (x += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
  }
  explicit Rng(uint64_t seed) { for (auto &v : s) v = splitmix(seed); }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {
    const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
  }
  double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  double uniform(double a, double b) { return a + (b - a) * uniform(); }
  int below(int n) { return (int)(uniform() * n) % n; }
};

struct Pose { double x, y, yaw; };
struct Discs { double fx, fy, rx, ry; };

constexpr double kRTurn = 3.0, kDeltat = 0.706;

Pose primitive(const Pose &s, int action) {
  const double r = kRTurn, d = kDeltat;
  const double dx = action == 0 ? r * d : r * std::sin(d);
  const double dy = action == 0 ? 0.0 : (action == 1 ? -r * (1 - std::cos(d)) : r * (1 - std::cos(d)));
  const double dyaw = action == 0 ? 0.0 : (action == 1 ? -d : d);
  const double c = std::cos(s.yaw), sn = std::sin(s.yaw);
  return Pose{s.x + dx * c - dy * sn, s.y + dx * sn + dy * c, s.yaw + dyaw};
}

struct Gen {
  double size, f2x, r2x, rv;
  std::vector<double> obs;                   // x, y, r
  std::vector<std::vector<Discs>> planned;   // per planned agent, per coarse state
  Discs discs(const Pose &s) const {
    const double c = std::cos(s.yaw), sn = std::sin(s.yaw);
    return Discs{s.x + f2x * c, s.y + f2x * sn, s.x + r2x * c, s.y + r2x * sn};
  }
  bool ok(const Pose &s, int step) const {
    const double margin = rv + 1.6, clear_o = rv + 0.35, clear_a = 2 * rv + 0.4;
    if (!(margin <= s.x && s.x <= size - margin && margin <= s.y && s.y <= size - margin)) return false;
    const Discs d = discs(s);
    const double lo = rv + 0.05, hi = size - rv - 0.05;
    if (!(d.fx > lo && d.fy > lo && d.rx > lo && d.ry > lo && d.fx < hi && d.fy < hi && d.rx < hi && d.ry < hi))
      return false;
    for (size_t o = 0; o + 2 < obs.size(); o += 3) {
      // the reference's static test is an axis-aligned square of half-size r + rv (corridor.cc:32-52)
      const double lim = obs[o + 2] + clear_o;
      if (std::fmax(std::fabs(d.fx - obs[o]), std::fabs(d.fy - obs[o + 1])) < lim) return false;
      if (std::fmax(std::fabs(d.rx - obs[o]), std::fabs(d.ry - obs[o + 1])) < lim) return false;
    }
    for (const auto &other : planned) {
      const Discs &q = other[step < (int)other.size() ? step : (int)other.size() - 1];
      const double c2 = clear_a * clear_a;
      auto d2 = [](double ax, double ay, double bx, double by) { return (ax - bx) * (ax - bx) + (ay - by) * (ay - by); };
      if (d2(d.fx, d.fy, q.fx, q.fy) < c2 || d2(d.fx, d.fy, q.rx, q.ry) < c2 || d2(d.rx, d.ry, q.fx, q.fy) < c2 ||
          d2(d.rx, d.ry, q.rx, q.ry) < c2)
        return false;
    }
    return true;
  }
};

}  // namespace

extern "C" {

// Horizon of an instance whose longest agent has `max_actions` coarse actions.
int synth_horizon(int max_actions) { return 3 * max_actions + 1; }

// One instance.  guess: [n_agents][6][nt_cap] is filled up to the instance's horizon *nt_out (planes of
// stride *nt_out, packed: guess[(a*6+k)* *nt_out + t]); obstacles: [n_obs][3].  Returns 0, or 1 when an
// agent could not be placed (lower the density).
int synth_instance(uint64_t seed, double size, int n_agents, int n_obs, int act_lo, int act_hi, double obs_radius,
                   double f2x, double r2x, double rv, double dt, double LF, double LB, double *guess, int nt_cap,
                   int *nt_out, double *obstacles, int *n_obs_out) {
  using namespace libMultiRobotPlanning;
  Rng rng(seed);
  Gen g{size, f2x, r2x, rv, {}, {}};
  int tries = 0;
  if (obs_radius < 0) {
    // room-like map (benchmark/room: 100x100, 130..300 discs of r = 0.5 on the integer grid, laid out as
    // wall segments with door gaps): horizontal / vertical walls on a 10 m lattice until n_obs discs exist
    const double r = -obs_radius;
    int guard = 0;
    while ((int)g.obs.size() / 3 < n_obs && guard++ < 400) {
      const bool horiz = rng.uniform() < 0.5;
      const int line = 10 * (1 + rng.below((int)(size / 10) - 1));        // wall position
      const int a0 = rng.below((int)size - 10), len = 10 + rng.below(31);  // start and length along the wall
      const int door = a0 + 3 + rng.below(len > 8 ? len - 6 : 1);          // a 6 m door gap
      for (int a = a0; a <= a0 + len && a <= (int)size && (int)g.obs.size() / 3 < n_obs; ++a) {
        if (a >= door && a < door + 6) continue;
        const double x = horiz ? a : line, y = horiz ? line : a;
        bool dup = false;
        for (size_t o = 0; o + 2 < g.obs.size(); o += 3) if (g.obs[o] == x && g.obs[o + 1] == y) { dup = true; break; }
        if (!dup) { g.obs.push_back(x); g.obs.push_back(y); g.obs.push_back(r); }
      }
    }
    n_obs = 0;  // (skip the uniform placement below)
  }
  while ((int)g.obs.size() / 3 < n_obs && tries < 100 * (n_obs > 0 ? n_obs : 1)) {  // generate_scenarios.py:83-102
    ++tries;
    const double x = rng.uniform(obs_radius, size - obs_radius), y = rng.uniform(obs_radius, size - obs_radius);
    bool free_ = true;
    for (size_t o = 0; o + 2 < g.obs.size(); o += 3)
      if ((x - g.obs[o]) * (x - g.obs[o]) + (y - g.obs[o + 1]) * (y - g.obs[o + 1]) <
          (obs_radius + g.obs[o + 2]) * (obs_radius + g.obs[o + 2])) { free_ = false; break; }
    if (free_) { g.obs.push_back(x); g.obs.push_back(y); g.obs.push_back(obs_radius); }
  }
  *n_obs_out = (int)g.obs.size() / 3;
  std::memcpy(obstacles, g.obs.data(), g.obs.size() * sizeof(double));
  const double margin = rv + 1.6;
  std::vector<CoarsePath> paths;
  for (int a = 0; a < n_agents; ++a) {
    bool placed = false;
    for (int attempt = 0; attempt < 400 && !placed; ++attempt) {
      Pose s{rng.uniform(margin, size - margin), rng.uniform(margin, size - margin), rng.uniform(-M_PI, M_PI)};
      if (!g.ok(s, 0)) continue;
      const int n_act = act_lo + rng.below(act_hi - act_lo + 1);
      std::vector<Pose> st{s};
      std::vector<int> acts;
      int cur = 0;
      bool alive = true;
      for (int k = 0; k < n_act && alive; ++k) {
        int order[3] = {0, 1, 2};
        for (int i = 2; i > 0; --i) { const int j = rng.below(i + 1); const int t = order[i]; order[i] = order[j]; order[j] = t; }
        if (rng.uniform() < 0.7) {  // keep the current primitive first
          int pos = 0;
          for (int i = 0; i < 3; ++i) if (order[i] == cur) pos = i;
          for (int i = pos; i > 0; --i) order[i] = order[i - 1];
          order[0] = cur;
        }
        bool moved = false;
        for (int i = 0; i < 3 && !moved; ++i) {
          const Pose nx = primitive(st.back(), order[i]);
          if (g.ok(nx, k + 1)) { st.push_back(nx); acts.push_back(order[i]); cur = order[i]; moved = true; }
        }
        if (!moved) {
          if (g.ok(st.back(), k + 1)) { st.push_back(st.back()); acts.push_back(6); }  // wait in place
          else alive = false;
        }
      }
      if (!alive && (int)acts.size() < act_lo / 2) continue;
      const int last = (int)st.size() - 1;
      if (!g.ok(st.back(), last + 5) || !g.ok(st.back(), last + 15) || !g.ok(st.back(), last + 40)) continue;
      CoarsePath cp;
      for (const Pose &q : st) cp.states.push_back(CoarseState{q.x, q.y, q.yaw});
      cp.actions = acts;
      paths.push_back(cp);
      std::vector<Discs> dd;
      for (const Pose &q : st) dd.push_back(g.discs(q));
      g.planned.push_back(dd);
      placed = true;
    }
    if (!placed) return 1;
  }
  std::vector<std::vector<OptimizeResult>> x0;
  InterpolateInitalGuess(paths, x0, nullptr, dt, LF, LB);
  const int nt = (int)x0[0].size();
  *nt_out = nt;
  if (nt > nt_cap) return 2;
  for (int a = 0; a < n_agents; ++a)
    for (int t = 0; t < nt; ++t) {
      const OptimizeResult &r = x0[a][t];
      double *p = guess + (size_t)a * 6 * nt;
      p[0 * nt + t] = r.x; p[1 * nt + t] = r.y; p[2 * nt + t] = r.yaw;
      p[3 * nt + t] = r.steer; p[4 * nt + t] = r.v; p[5 * nt + t] = r.d_steer;
    }
  return 0;
}

// InterpolateInitalGuess (include/csdo/initial_guess.h) for Na coarse paths given as flat arrays: states
// [sum n_states][3], actions [sum n_states] (one unused slot per agent at its end), goals [Na][3] or NULL.
// out: Na x 6 planes of nt_cap; returns the horizon or -1.
int interp_paths(int Na, const int *n_states, const double *states, const signed char *actions, const double *goals,
                 double dt, double LF, double LB, int nt_cap, double *out) {
  using namespace libMultiRobotPlanning;
  std::vector<CoarsePath> paths(Na);
  std::vector<CoarseState> gl(Na);
  size_t so = 0;
  for (int a = 0; a < Na; ++a) {
    for (int i = 0; i < n_states[a]; ++i, ++so) {
      paths[a].states.push_back(CoarseState{states[3 * so], states[3 * so + 1], states[3 * so + 2]});
      if (i + 1 < n_states[a]) paths[a].actions.push_back((int)actions[so]);
    }
    if (goals) gl[a] = CoarseState{goals[3 * a], goals[3 * a + 1], goals[3 * a + 2]};
  }
  std::vector<std::vector<OptimizeResult>> x0;
  InterpolateInitalGuess(paths, x0, goals ? &gl : nullptr, dt, LF, LB);
  const int nt = (int)x0[0].size();
  if (nt > nt_cap) return -1;
  for (int a = 0; a < Na; ++a)
    for (int t = 0; t < nt; ++t) {
      const OptimizeResult &r = x0[a][t];
      double *p = out + (size_t)a * 6 * nt_cap;
      p[t] = r.x; p[nt_cap + t] = r.y; p[2 * nt_cap + t] = r.yaw; p[3 * nt_cap + t] = r.steer; p[4 * nt_cap + t] = r.v;
      p[5 * nt_cap + t] = r.d_steer;
    }
  return nt;
}

}  // extern "C"
