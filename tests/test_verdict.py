"""CPU tests of the verdict / output-format restatement against golden vectors
produced by the reference's own scripts/collision_detection.py (tests/golden/make_verdict_golden.py)."""
import os

import numpy as np

from csdotrajectoryplanning_b200 import verdict as V

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "verdict_golden.npz")


def test_rect_rect_matches_reference_vectors():
    g = np.load(GOLD)["rect_rect"]
    assert g.shape[0] == 400 and 100 < g[:, -1].sum() < 300
    for row in g:
        got = V.collision_rect_and_rect(V._rect(row[0:3]), V._rect(row[3:6]))
        assert got == bool(row[6])


def test_circle_rect_matches_reference_vectors():
    g = np.load(GOLD)["circle_rect"]
    for row in g:
        got = V.collision_circle_and_rect(tuple(row[3:6]), V._rect(row[0:3]))
        assert got == bool(row[6])


def test_verdict_loop_and_success_rule(oracle, params, small_batch):
    res, _ = oracle.refine(params, small_batch, linsys=1, nthreads=2)
    b = small_batch
    for i in range(b.n_inst):
        a0, a1 = int(b.inst_agent_ptr[i]), int(b.inst_agent_ptr[i + 1])
        trajs = [V.rounded_solution(res.agent_traj(b, a)) for a in range(a0, a1)]
        obs = b.obs[3 * b.obs_ptr[i]:3 * b.obs_ptr[i + 1]].reshape(-1, 3)
        inter, static = V.verdict(trajs, obs)
        # the refined synthetic plans are collision free and count as solved
        assert inter == [] and static == []
        assert V.success(res.inst_status[i])
    # two overlapping parked cars and a car on an obstacle are reported
    t1 = np.array([[10.0] * 3, [10.0] * 3, [0.0] * 3])
    t2 = np.array([[11.0] * 3, [10.5] * 3, [0.3] * 3])
    inter, static = V.verdict([t1, t2], np.array([[10.5, 10.0, 0.8]]))
    assert len(inter) == 3 and inter[0] == (0, 0, 1) and len(static) == 6
    assert V.success(-2) and V.success(2) and not V.success(-3) and not V.success(3)


def test_three_decimal_format():
    tr = np.array([[1.23449, 2.0005], [0.0, -0.00049], [3.14159, -3.14159]])
    r = V.rounded_solution(tr)
    assert r.tolist() == [[1.234, 2.0], [0.0, -0.0], [3.142, -3.142]] or r[0, 1] in (2.0, 2.001)


def test_prefilters_do_not_change_the_verdict():
    """verdict() skips far-apart pairs before the exact tests: same lists as testing every pair."""
    from csdotrajectoryplanning_b200 import verdict as V
    rng = np.random.default_rng(11)
    for trial in range(6):
        na, nt, no = 7, 5, 9
        size = (12.0, 25.0, 60.0)[trial % 3]                     # crowded ... sparse
        trajs = [np.vstack([rng.uniform(0, size, (2, nt)), rng.uniform(-3.2, 3.2, (1, nt))]) for _ in range(na)]
        obs = np.column_stack([rng.uniform(0, size, (no, 2)), rng.uniform(0.3, 2.5, no)])
        if trial == 5:
            obs = obs[:, :2]                                     # 2-element obstacles: radius 1
        inter, static = V.verdict(trajs, obs)
        bi, bs = [], []
        for f in range(nt):
            rects = [V._rect(t[:3, f]) for t in trajs]
            for ai in range(na):
                for aj in range(ai + 1, na):
                    if V.collision_rect_and_rect(rects[ai], rects[aj]):
                        bi.append((f, ai, aj))
            for a in range(na):
                for oi, o in enumerate(obs):
                    if V.collision_circle_and_rect((o[0], o[1], o[2] if len(o) == 3 else V.OBS_RADIUS_VIS), rects[a]):
                        bs.append((f, a, oi))
        assert inter == bi and static == bs
        if trial % 3 == 0:
            assert bi and bs                                      # the crowded scenes do collide
