"""Host-side input construction for the DSQP refine stage.

* :func:`interpolate_initial_guess` -- the reference's ``InterpolateInitalGuess``
  chain (sqp/inter_agent_cons.cc:143-411): coarse (state, action) path ->
  ``x0_bar`` planes x,y,yaw,steer,v,w.  Host logic, double precision.
* :func:`synthetic_instance` / :func:`synthetic_batch` -- seeded stand-ins for
  the PBS + Hybrid-A* coarse guess (pbs/, hybrid_a_star/ stay on the host in
  the reference and are out of scope here): each agent follows a random
  sequence of the planner's own motion primitives (common/motion_planning.cc
  :47-51, step r*deltat = 2.118 m, turn deltat = 0.706 rad at radius r = 3)
  that avoids the obstacles, the map border and the already planned agents,
  i.e. it has the shape of a priority-based plan.
* :func:`load_scenario_yaml` -- benchmark YAML (hybrid_a_star/Instance.cc:6-63).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .batch import Batch, Instance, pack_instances
from .params import CsdoParams, default_params

R_TURN = 3.0      # config.yaml r
DELTAT = 0.706    # config.yaml deltat
OBS_RADIUS = 0.8  # config.yaml obsRadius


def _normalize_angle_abs_in_pi(x: float) -> float:
    """motion_planning.h:70-75 -- note the float return type."""
    x = math.fmod(x + math.pi, 2 * math.pi)
    if x < 0:
        x += 2 * math.pi
    return float(np.float32(x - math.pi))


def interpolate_path(states: np.ndarray, actions: Sequence[int], n: int,
                     r_const: float) -> Tuple[np.ndarray, List[int]]:
    """interpolateXYYaw + action_sample (inter_agent_cons.cc:194-311)."""
    out = [tuple(states[0])]
    acts: List[int] = []
    s = tuple(states[0])
    for i, action in enumerate(actions):
        s0, s1 = s, tuple(states[i + 1])
        acts.extend([int(action)] * (n + 1))
        if action == 6:
            out.extend([s0] * (n + 1))
        else:
            r = r_const
            if action in (0, 3):
                deltat = math.sqrt((s1[0] - s0[0]) ** 2 + (s1[1] - s0[1]) ** 2) / r_const
            else:
                deltat = _normalize_angle_abs_in_pi(s1[2] - s0[2])
                d = math.sqrt((s1[0] - s0[0]) ** 2 + (s1[1] - s0[1]) ** 2)
                r = d / (2.0 * math.sin(abs(deltat) / 2.0))
            da = abs(deltat) / float(n + 1)
            dx = (r * da, r * math.sin(da), r * math.sin(da),
                  -r * da, -r * math.sin(da), -r * math.sin(da))[action]
            dy = (0.0, -r * (1 - math.cos(da)), r * (1 - math.cos(da)),
                  0.0, -r * (1 - math.cos(da)), r * (1 - math.cos(da)))[action]
            dyaw = (0.0, -da, da, 0.0, da, -da)[action]
            c = list(s0)
            for _ in range(n):
                xs = c[0] + dx * math.cos(c[2]) - dy * math.sin(c[2])
                ys = c[1] + dx * math.sin(c[2]) + dy * math.cos(c[2])
                c = [xs, ys, c[2] + dyaw]
                out.append(tuple(c))
            out.append((s1[0], s1[1], (0.0 if action in (0, 3) else deltat) + s0[2]))
        s = out[-1]
    return np.asarray(out, np.float64), acts


def interpolate_initial_guess(paths: Sequence[Tuple[np.ndarray, Sequence[int]]],
                              goals: Optional[np.ndarray], params: CsdoParams,
                              num_interpolation: int = 2, r_const: float = R_TURN) -> np.ndarray:
    """InterpolateInitalGuess (inter_agent_cons.cc:143-157) -> (Na, 6, Nt)."""
    fine = []
    for a, (states, actions) in enumerate(paths):
        st = np.array(states, np.float64, copy=True)
        if goals is not None:
            st[-1] = goals[a]                       # :149-151
        fine.append(interpolate_path(st, actions, num_interpolation, r_const))
    nt = max(f[0].shape[0] for f in fine)
    phi = float(np.arctan(np.float32((np.float32(params.LF) - np.float32(params.LB))
                                     / np.float32(r_const))))  # std::atan(float), :352-353
    dt = params.dt
    g = np.zeros((len(fine), 6, nt))
    for a, (st, acts) in enumerate(fine):
        ns = st.shape[0]
        g[a, 0, :ns], g[a, 1, :ns], g[a, 2, :ns] = st[:, 0], st[:, 1], st[:, 2]
        g[a, 0, ns:], g[a, 1, ns:], g[a, 2, ns:] = st[-1, 0], st[-1, 1], st[-1, 2]
        for i in range(1, ns):
            act = acts[i - 1]
            g[a, 3, i] = 0.0 if act in (0, 3, 6) else (-phi if act in (1, 4) else phi)
        x, y, yaw, steer = g[a, 0], g[a, 1], g[a, 2], g[a, 3]
        g[a, 4, :ns - 1] = ((x[1:ns] - x[:ns - 1]) / dt) * np.cos(yaw[:ns - 1]) + \
                           ((y[1:ns] - y[:ns - 1]) / dt) * np.sin(yaw[:ns - 1])
        g[a, 5, :ns - 1] = (steer[1:ns] - steer[:ns - 1]) / dt
    return g


# --------------------------------------------------------------------------
# synthetic coarse plans
def _primitive(s: np.ndarray, action: int) -> np.ndarray:
    """One planner step (Constants::dx/dy/dyaw, motion_planning.cc:96-108)."""
    r, d = R_TURN, DELTAT
    dx = (r * d, r * math.sin(d), r * math.sin(d))[action]
    dy = (0.0, -r * (1 - math.cos(d)), r * (1 - math.cos(d)))[action]
    dyaw = (0.0, -d, d)[action]
    c, sn = math.cos(s[2]), math.sin(s[2])
    return np.array([s[0] + dx * c - dy * sn, s[1] + dx * sn + dy * c, s[2] + dyaw])


def _discs(s: np.ndarray, p: CsdoParams) -> np.ndarray:
    c, sn = math.cos(s[2]), math.sin(s[2])
    return np.array([[s[0] + p.f2x * c, s[1] + p.f2x * sn], [s[0] + p.r2x * c, s[1] + p.r2x * sn]])


def synthetic_obstacles(rng: np.random.Generator, size: float, n_obs: int,
                        radius: float = OBS_RADIUS) -> np.ndarray:
    """Uniform discs rejecting overlaps (scripts/generate_scenarios.py:83-102)."""
    obs: List[List[float]] = []
    tries = 0
    while len(obs) < n_obs and tries < 100 * max(n_obs, 1):
        tries += 1
        x, y = rng.uniform(radius, size - radius, 2)
        if all((x - o[0]) ** 2 + (y - o[1]) ** 2 >= (radius + o[2]) ** 2 for o in obs):
            obs.append([x, y, radius])
    return np.asarray(obs, np.float64).reshape(-1, 3)


def synthetic_instance(seed: int, size: float, n_agents: int, n_obs: int,
                       n_actions: Tuple[int, int], params: Optional[CsdoParams] = None,
                       name: str = "") -> Instance:
    """One seeded instance with a priority-style collision-free coarse plan."""
    p = params or default_params()
    rng = np.random.default_rng(seed)
    obs = synthetic_obstacles(rng, size, n_obs)
    margin = p.rv + 1.6           # keep both discs inside [rv, size-rv]
    clear_o = p.rv + 0.35         # disc-centre clearance to an obstacle edge
    clear_a = 2 * p.rv + 0.4      # disc-centre clearance between two agents
    planned: List[np.ndarray] = []   # per agent (n_states, 2, 2) disc centres

    def state_ok(s: np.ndarray, step: int) -> bool:
        if not (margin <= s[0] <= size - margin and margin <= s[1] <= size - margin):
            return False
        d = _discs(s, p)
        if not (np.all(d > p.rv + 0.05) and np.all(d < size - p.rv - 0.05)):
            return False
        if obs.shape[0]:
            # the reference's static test is an axis-aligned square of half-size r + rv
            # around each obstacle (corridor.cc:32-52), hence the Chebyshev distance
            dd = np.maximum(np.abs(d[:, None, 0] - obs[None, :, 0]), np.abs(d[:, None, 1] - obs[None, :, 1]))
            if np.any(dd < obs[None, :, 2] + clear_o):
                return False
        for other in planned:
            o = other[min(step, other.shape[0] - 1)]
            dd = np.hypot(d[:, None, 0] - o[None, :, 0], d[:, None, 1] - o[None, :, 1])
            if np.any(dd < clear_a):
                return False
        return True

    paths = []
    for _ in range(n_agents):
        for attempt in range(200):
            s = np.array([rng.uniform(margin, size - margin), rng.uniform(margin, size - margin),
                          rng.uniform(-math.pi, math.pi)])
            if not state_ok(s, 0):
                continue
            # a parked agent must not sit on an earlier agent's future path either
            n_act = int(rng.integers(n_actions[0], n_actions[1] + 1))
            states, acts, cur, ok = [s], [], 0, True
            for k in range(n_act):
                order = [cur] + [a for a in rng.permutation(3) if a != cur] \
                    if rng.random() < 0.7 else list(rng.permutation(3))
                for a in order:
                    nxt = _primitive(states[-1], int(a))
                    if state_ok(nxt, k + 1):
                        states.append(nxt); acts.append(int(a)); cur = int(a)
                        break
                else:
                    # blocked: wait in place (planner action 6) if that is collision free
                    if state_ok(states[-1], k + 1):
                        states.append(states[-1].copy()); acts.append(6)
                    else:
                        ok = False
                        break
            if not ok and len(acts) < n_actions[0] // 2:
                continue
            # the final pose is held for the rest of the horizon: check it stays clear
            last = len(states) - 1
            if any(not state_ok(states[-1], last + j) for j in (5, 15, 40)):
                continue
            paths.append((np.asarray(states), acts))
            planned.append(np.stack([_discs(st, p) for st in states]))
            break
        else:
            raise RuntimeError("could not place an agent; lower the density")
    guess = interpolate_initial_guess(paths, None, p)
    return Instance(guess, size, size, obs, None, None, name or f"synthetic_{seed}")


def synthetic_batch(shapes: Sequence[Tuple[float, int, int, Tuple[int, int]]], per_shape: int,
                    seed: int = 1234, params: Optional[CsdoParams] = None) -> List[Instance]:
    """shapes: (map size, n_agents, n_obstacles, (min,max) coarse actions)."""
    out = []
    k = 0
    for (size, na, no, nact) in shapes:
        for j in range(per_shape):
            out.append(synthetic_instance(seed + k, size, na, no, nact, params,
                                          f"map{int(size)}_a{na}_o{no}_ex{j}"))
            k += 1
    return out


MAP50_SWEEP = [(50.0, na, no, (12, 30)) for na in (5, 10, 15, 20, 25) for no in (0, 25)]
MAP100_A100 = [(100.0, 100, 50, (30, 60))]


def load_scenario_yaml(path: str):
    """Benchmark scenario (Instance.cc:6-63): dims, obstacles (x,y[,r]), starts, goals."""
    import yaml
    with open(path) as f:
        doc = yaml.safe_load(f)
    dims = doc["map"]["dimensions"]
    obs = []
    for o in (doc["map"].get("obstacles") or []):
        obs.append([float(o[0]), float(o[1]), float(o[2]) if len(o) > 2 else OBS_RADIUS])
    starts = np.asarray([a["start"] for a in doc["agents"]], np.float64)
    goals = np.asarray([a["goal"] for a in doc["agents"]], np.float64)
    return float(dims[0]), float(dims[1]), np.asarray(obs, np.float64).reshape(-1, 3), starts, goals
