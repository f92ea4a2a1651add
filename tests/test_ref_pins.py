"""CPU tests that pin the oracle and the host logic against the REFERENCE'S OWN code.

tests/golden/ref_pins.npz was minted by tests/golden/make_ref_golden.py from oracle/_ref/libcsdo_ref.so,
i.e. /root/reference/sqp/corridor.cc + sqp/inter_agent_cons.cc compiled unmodified (`make -C oracle ref`).
Rows pinned: a2 generateBox/calcCorridors, a14/a15 pairs + planes, f2
InterpolateInitalGuess, f3 dumpSolutions.  Everything is bit-for-bit.  The live tests at the end run only
where the reference build exists (this container); the fixture tests run everywhere.
"""
import os
import subprocess

import numpy as np
import pytest

from csdotrajectoryplanning_b200.output import SolutionStatistics, format_solutions
from csdotrajectoryplanning_b200.scenario import interpolate_initial_guess

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def pins():
    return np.load(os.path.join(HERE, "golden", "ref_pins.npz"))


def _paths(pins):
    ns, st, ac = pins["path_ns"], pins["path_states"], pins["path_actions"]
    paths, so, ao = [], 0, 0
    for n in ns:
        paths.append((st[so:so + n].copy(), ac[ao:ao + n - 1].copy()))
        so += n; ao += n - 1
    return paths


def test_oracle_generate_box_matches_reference(pins, oracle, params):
    """a2: 400 reference boxes (random points incl. out-of-map and in-collision starts, 0..79 obstacles of
    r = 0.5 / 0.8); obstacles are fed in the iteration order of the reference's unordered_set."""
    ptr, obs, order = pins["box_obs_ptr"], pins["box_obs"], pins["box_order"]
    seen = set()
    for k in range(pins["box_xy"].shape[0]):
        o = obs[ptr[k]:ptr[k + 1]][order[ptr[k]:ptr[k + 1]]]
        box, st = oracle.generate_box(params, float(pins["box_size"][k]), float(pins["box_size"][k]),
                                      float(pins["box_xy"][k, 0]), float(pins["box_xy"][k, 1]), o)
        assert np.array_equal(box, pins["box_out"][k]), k
        assert np.array_equal(st, pins["box_status"][k]), k
        seen.add(tuple(int(v) for v in pins["box_status"][k]))
    assert {(1, 0), (1, 1), (1, 2)} <= seen          # legal, out-of-map and in-collision starts all occur


def test_python_and_oracle_initial_guess_match_reference(pins, oracle, params):
    """f2: InterpolateInitalGuess of the reference on paths with all 7 planner actions and snapped goals."""
    paths = _paths(pins)
    want = pins["guess"]
    got = interpolate_initial_guess(paths, pins["path_goals"], params)
    assert got.shape == want.shape and np.array_equal(got, want)
    for a, (st, ac) in enumerate(paths):
        g, _ = oracle.interpolate_guess(st, ac, pins["path_goals"][a], 2, params.dt, 3.0, params.LF, params.LB,
                                        want.shape[2])
        assert np.array_equal(g, want[a]), a


def test_cpp_initial_guess_matches_reference(pins, params):
    subprocess.run(["make", "-C", os.path.join(HERE, "cpp"), "test_initial_guess"], check=True, capture_output=True)
    paths, goals, want = _paths(pins), pins["path_goals"], pins["guess"]
    lines = [f"{len(paths)} {params.dt!r} {params.LF!r} {params.LB!r}"]
    for (st, ac), g in zip(paths, goals):
        lines.append(str(len(ac)))
        lines += [" ".join(repr(float(v)) for v in s) for s in st]
        lines.append(" ".join(str(int(a)) for a in ac))
        lines.append(" ".join(repr(float(v)) for v in g))
    out = subprocess.run([os.path.join(HERE, "cpp", "test_initial_guess")], input="\n".join(lines) + "\n",
                         capture_output=True, text=True, check=True).stdout.split()
    nt = int(out[0])
    got = np.array([float(v) for v in out[1:]]).reshape(len(paths), nt, 6)
    assert nt == want.shape[2]
    assert np.array_equal(got.transpose(0, 2, 1), want)


def test_oracle_planes_match_reference(pins, oracle, params):
    """a14/a15: pair list (through the per-agent plane times) and the 12 coefficients, bit for bit."""
    pt, pabc, legal = oracle.instance_planes(params, pins["guess"])
    assert np.array_equal(np.asarray([len(t) for t in pt]), pins["plane_cnt"])
    assert np.array_equal(np.concatenate(pt), pins["plane_t"])
    assert np.array_equal(np.concatenate(pabc), pins["plane_abc"])
    assert legal == bool(pins["inter_legal"][0])
    assert int(pins["plane_cnt"].sum()) == 2 * int(pins["n_pairs"][0])


def test_oracle_corridors_match_reference(pins, oracle, params):
    """a2: calcCorridors (float disc centres through State) for a 9-agent instance with 25 obstacles."""
    g, obs = pins["guess"], pins["corr_obs"][pins["corr_order"]]
    legal_all = True
    for a in range(g.shape[0]):
        corr, _, legal = oracle.agent_corridors(params, g[a, 0], g[a, 1], g[a, 2], 50.0, 50.0, obs, False)
        assert np.array_equal(corr, pins["corr"][a]), a
        legal_all &= legal
    assert legal_all == bool(pins["static_legal"][0])


def test_writers_match_reference_dump_solutions(pins, tmp_path):
    """f3: the Python and the C++ writer produce the reference's dumpSolutions bytes."""
    want = bytes(pins["dump_text"]).decode()
    s = pins["dump_stat"]
    stat = SolutionStatistics(*[float(v) for v in s[:8]], int(s[8]), int(s[9]))
    assert format_solutions(pins["guess"][:3], stat) == want
    subprocess.run(["make", "-C", os.path.join(HERE, "cpp"), "test_solution_io"], check=True, capture_output=True)
    g = pins["guess"][:3]
    na, _, nt = g.shape
    lines = [f"{na} {nt}", " ".join(repr(float(v)) for v in s)]
    lines += [" ".join(repr(float(v)) for v in g[a, :, t]) for a in range(na) for t in range(nt)]
    out = str(tmp_path / "cpp.yaml")
    subprocess.run([os.path.join(HERE, "cpp", "test_solution_io"), out, "full"], input="\n".join(lines) + "\n", text=True,
                   check=True)
    assert open(out).read() == want


# ---- live comparisons (need oracle/_ref, i.e. /root/reference or a prebuilt copy) ------------------
def _ref():
    from oracle import ref as R
    if not R.available():
        pytest.skip("oracle/_ref not built here (no /root/reference)")
    return R


def test_live_thousand_random_boxes(oracle, params):
    R = _ref()
    rng = np.random.default_rng(7)
    for k in range(1000):
        no = int(rng.integers(0, 300)) if k % 10 == 0 else int(rng.integers(0, 60))
        size = 100.0 if k % 3 == 0 else 50.0
        obs = np.column_stack([rng.uniform(0, size, no), rng.uniform(0, size, no), rng.choice([0.5, 0.8, 1.5], no)]) \
            if no else np.zeros((0, 3))
        x, y = rng.uniform(-1, size + 1, 2)
        b2, s2 = R.generate_box(size, size, float(x), float(y), obs)
        o = obs[R.obstacle_order(obs)] if no else obs
        b1, s1 = oracle.generate_box(params, size, size, float(x), float(y), o)
        assert np.array_equal(b1, b2) and np.array_equal(s1, s2), k


def test_live_planes_and_corridors_on_synthetic_instance(oracle, params):
    R = _ref()
    from csdotrajectoryplanning_b200.scenario import synthetic_instance
    ins = synthetic_instance(31, 50.0, 10, 25, (10, 18), params)
    pt1, pa1, l1 = oracle.instance_planes(params, ins.guess)
    pt2, pa2, l2, _ = R.instance_planes(ins.guess, params.r_trust)
    assert l1 == l2 and all(np.array_equal(a, b) for a, b in zip(pt1, pt2))
    assert all(np.array_equal(a, b) for a, b in zip(pa1, pa2))
    corr, _ = R.calc_corridors(ins.guess, 50.0, 50.0, ins.obstacles)
    obs = ins.obstacles[R.obstacle_order(ins.obstacles)]
    for a in range(ins.n_agents):
        c1, _, _ = oracle.agent_corridors(params, ins.guess[a, 0], ins.guess[a, 1], ins.guess[a, 2], 50.0, 50.0, obs, False)
        assert np.array_equal(c1, corr[a])
