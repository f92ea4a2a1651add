"""CPU test of the C++ interpolated-initial-guess header (include/csdo/initial_guess.h):
bit-identical to the Python host logic (scenario.interpolate_initial_guess), which the oracle tests check
against the C restatement of inter_agent_cons.cc:143-411."""
import os
import subprocess

import numpy as np

import math

from csdotrajectoryplanning_b200.scenario import DELTAT, R_TURN, interpolate_initial_guess

HERE = os.path.dirname(os.path.abspath(__file__))


def _primitive(s, action):
    """The planner's seven motion primitives (common/motion_planning.cc:47-51, 96-108), 6 = wait."""
    if action == 6:
        return s.copy()
    r, d = R_TURN, DELTAT
    sx, cy = r * math.sin(d), r * (1 - math.cos(d))
    dx = (r * d, sx, sx, -r * d, -sx, -sx)[action]
    dy = (0.0, -cy, cy, 0.0, -cy, cy)[action]
    dyaw = (0.0, -d, d, 0.0, d, -d)[action]
    c, sn = math.cos(s[2]), math.sin(s[2])
    return np.array([s[0] + dx * c - dy * sn, s[1] + dx * sn + dy * c, s[2] + dyaw])


def _random_paths(rng, na):
    paths, goals = [], []
    for _ in range(na):
        n = int(rng.integers(3, 9))
        s = np.array([rng.uniform(10, 40), rng.uniform(10, 40), rng.uniform(-3, 3)])
        states, actions = [s.copy()], []
        for _ in range(n):
            a = int(rng.integers(0, 7))
            s = _primitive(s, a)
            states.append(s.copy()); actions.append(a)
        paths.append((np.array(states), actions))
        goals.append(states[-1] + rng.normal(0, 0.01, 3))   # the planner's goal snap (:149-151)
    return paths, np.array(goals)


def test_cpp_initial_guess_matches_host_logic(params):
    subprocess.run(["make", "-C", os.path.join(HERE, "cpp"), "test_initial_guess"], check=True, capture_output=True)
    rng = np.random.default_rng(5)
    for na in (1, 4, 7):
        paths, goals = _random_paths(rng, na)
        want = interpolate_initial_guess(paths, goals, params)
        lines = [f"{na} {params.dt!r} {params.LF!r} {params.LB!r}"]
        for (st, ac), g in zip(paths, goals):
            lines.append(str(len(ac)))
            lines += [" ".join(repr(float(v)) for v in s) for s in st]
            lines.append(" ".join(str(a) for a in ac))
            lines.append(" ".join(repr(float(v)) for v in g))
        out = subprocess.run([os.path.join(HERE, "cpp", "test_initial_guess")], input="\n".join(lines) + "\n",
                             capture_output=True, text=True, check=True).stdout.split()
        nt = int(out[0])
        got = np.array([float(v) for v in out[1:]]).reshape(na, nt, 6)
        assert nt == want.shape[2]
        for k, col in enumerate((0, 1, 2, 3, 4, 5)):       # x, y, yaw, steer, v, w
            assert np.array_equal(got[:, :, col], want[:, k, :]), k
