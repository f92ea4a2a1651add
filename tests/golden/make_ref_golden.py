"""Mint tests/golden/ref_pins.npz from the REFERENCE'S OWN code (test infrastructure).

oracle/_ref/libcsdo_ref.so is /root/reference/sqp/corridor.cc + sqp/inter_agent_cons.cc compiled
unmodified (`make -C oracle ref`).  Rows pinned: a2 corridors (generateBox,
calcCorridors), a14/a15 neighbour pairs + planes, f2 InterpolateInitalGuess, f3 dumpSolutions, and the
iteration order of std::unordered_set<Location>.  Runs only where /root/reference exists; the fixture
travels.  Regenerate with:  python tests/golden/make_ref_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from csdotrajectoryplanning_b200 import default_params  # noqa: E402
from csdotrajectoryplanning_b200.scenario import _primitive  # noqa: E402
from oracle import ref as R  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def coarse_paths(rng, na, size):
    """random coarse (state, action) paths incl. reverse primitives and waits (all 7 planner actions)"""
    paths, goals = [], []
    for _ in range(na):
        s = np.array([rng.uniform(8, size - 8), rng.uniform(8, size - 8), rng.uniform(-np.pi, np.pi)])
        n = int(rng.integers(5, 14))
        st, ac = [s], []
        for _ in range(n):
            a = int(rng.choice([0, 0, 1, 2, 3, 4, 5, 6]))
            if a == 6:
                nx = st[-1].copy()
            elif a < 3:
                nx = _primitive(st[-1], a)
            else:   # reverse primitives: Constants::dx/dy/dyaw entries 3..5 (motion_planning.cc:96-108)
                r, d = 3.0, 0.706
                dx = (-r * d, -r * np.sin(d), -r * np.sin(d))[a - 3]
                dy = (0.0, -r * (1 - np.cos(d)), r * (1 - np.cos(d)))[a - 3]
                dyaw = (0.0, d, -d)[a - 3]
                c, sn = np.cos(st[-1][2]), np.sin(st[-1][2])
                nx = np.array([st[-1][0] + dx * c - dy * sn, st[-1][1] + dx * sn + dy * c, st[-1][2] + dyaw])
            st.append(nx); ac.append(a)
        if ac[-1] == 6:            # a path never ends on a wait in the reference's planner
            ac[-1] = 0; st[-1] = _primitive(st[-2], 0)
        paths.append((np.asarray(st), np.asarray(ac, np.int32)))
        goals.append(st[-1] + np.array([0.03, -0.02, 0.01]))   # the goal differs slightly from the last state
    return paths, np.asarray(goals)


def main():
    p = default_params()
    rng = np.random.default_rng(20261017)
    out = {}
    # ---- a2: generateBox on random points / obstacle sets (incl. out-of-map, in-collision, room-like r=0.5) ----
    cases = []
    for k in range(400):
        no = int(rng.integers(0, 80))
        size = float(rng.choice([50.0, 100.0]))
        rad = float(rng.choice([0.5, 0.8]))
        obs = np.column_stack([rng.uniform(0, size, no), rng.uniform(0, size, no), np.full(no, rad)]) if no \
            else np.zeros((0, 3))
        x, y = rng.uniform(-1, size + 1, 2)
        if no and k % 5 == 0:       # start inside an obstacle's inflated square
            j = int(rng.integers(no)); x, y = obs[j, 0] + rng.uniform(-1.5, 1.5), obs[j, 1] + rng.uniform(-1.5, 1.5)
        box, st = R.generate_box(size, size, float(x), float(y), obs)
        order = R.obstacle_order(obs) if no else np.zeros(0, np.int32)
        cases.append((size, x, y, obs, order, box, st))
    out["box_size"] = np.asarray([c[0] for c in cases]); out["box_xy"] = np.asarray([[c[1], c[2]] for c in cases])
    out["box_obs_ptr"] = np.cumsum([0] + [c[3].shape[0] for c in cases]).astype(np.int32)
    out["box_obs"] = np.concatenate([c[3] for c in cases]); out["box_order"] = np.concatenate([c[4] for c in cases]).astype(np.int32)
    out["box_out"] = np.asarray([c[5] for c in cases]); out["box_status"] = np.asarray([c[6] for c in cases])
    # ---- f2: InterpolateInitalGuess; a14/a15: pairs + planes; a2: calcCorridors on the same instance ----
    paths, goals = coarse_paths(rng, 9, 50.0)
    guess = R.interpolate_guess(paths, goals, 2, p.dt)
    out["path_ns"] = np.asarray([len(s) for s, _ in paths], np.int32)
    out["path_states"] = np.concatenate([s for s, _ in paths]); out["path_actions"] = np.concatenate([a for _, a in paths])
    out["path_goals"] = goals; out["guess"] = guess
    pt, pabc, legal, npairs = R.instance_planes(guess, p.r_trust)
    out["plane_cnt"] = np.asarray([len(t) for t in pt], np.int32); out["plane_t"] = np.concatenate(pt)
    out["plane_abc"] = np.concatenate(pabc); out["inter_legal"] = np.asarray([legal]); out["n_pairs"] = np.asarray([npairs])
    obs = np.column_stack([rng.uniform(3, 47, 25), rng.uniform(3, 47, 25), np.full(25, 0.8)])
    corr, static_legal = R.calc_corridors(guess, 50.0, 50.0, obs)
    out["corr_obs"] = obs; out["corr_order"] = R.obstacle_order(obs); out["corr"] = corr
    out["static_legal"] = np.asarray([static_legal])
    # ---- f3: dumpSolutions text ----
    stat = np.asarray([-1, 37.5, 301.25, 12.3456, 10.0, 0.0456, 2.3, 0.125, 2, -2], np.float64)
    txt = R.dump_solutions(guess[:3], stat)
    out["dump_stat"] = stat; out["dump_text"] = np.frombuffer(txt.encode(), np.uint8)
    np.savez_compressed(os.path.join(HERE, "ref_pins.npz"), **out)
    print("ref_pins.npz written:", len(cases), "boxes,", guess.shape, "guess,", int(out["plane_cnt"].sum()), "planes,",
          len(txt), "bytes of dumpSolutions text; inter_legal", legal, "static_legal", static_legal)


if __name__ == "__main__":
    main()
