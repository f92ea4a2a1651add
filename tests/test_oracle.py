"""CPU tests: the oracle against known answers, invariants and golden fixtures.

PARITY UNPINNED with respect to the reference tree (it has no tests, golden
vectors or fixtures for this path).  The pins used here:
  * the published OSQP demo problem and its printed solver log (OSQP docs,
    "Demo"/"Setup and solve" example run with alpha=1.0): iteration 1
    objective -4.9384e-03, pri res 1.00e+00, dua res 2.00e+02, rho 1.00e-01;
    iteration 50 objective 1.8800e+00, pri res 1.91e-07, dua res 7.50e-07,
    rho 1.38e+00, status solved -- reproduced digit for digit;
  * KKT optimality of converged solutions, checked with numpy only;
  * the reference's own in-code diagnostics turned into assertions
    (dsqp_solver.cc:726-729, 775-782, 954-963, 991-992, 1125-1128);
  * fixtures minted by tests/golden/make_golden.py.
"""
import os

import numpy as np
import pytest

from csdotrajectoryplanning_b200 import default_params, pack_instances
from csdotrajectoryplanning_b200.scenario import synthetic_instance

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# OSQP documentation demo: P=[[4,1],[1,2]], q=[1,1], A=[[1,1],[1,0],[0,1]], l=[1,0,0], u=[1,.7,.7]
DEMO = dict(n=2, m=3, Pp=[0, 1, 3], Pi=[0, 0, 1], Px=[4., 1., 2.], q=[1., 1.],
            Ap=[0, 2, 4], Ai=[0, 1, 0, 2], Ax=[1., 1., 1., 1.], l=[1., 0., 0.], u=[1., .7, .7])


def _demo(oracle, p, **kw):
    d = DEMO
    return oracle.osqp_solve(p, d["n"], d["m"], d["Pp"], d["Pi"], d["Px"], d["q"], d["Ap"], d["Ai"],
                             d["Ax"], d["l"], d["u"], **kw)


def test_osqp_demo_published_log(oracle):
    p = default_params()
    p.alpha = 1.0  # the documentation example changes alpha to 1.0
    r1 = _demo(oracle, p, max_iter=1)
    assert "%.4e" % r1["obj"] == "-4.9384e-03"
    assert "%.2e" % r1["pri_res"] == "1.00e+00"
    assert "%.2e" % r1["dua_res"] == "2.00e+02"
    assert "%.2e" % r1["rho"] == "1.00e-01"
    r = _demo(oracle, p, max_iter=4000)
    assert r["status"] == 1 and r["iters"] == 50 and r["n_factor"] == 2
    assert "%.4e" % r["obj"] == "1.8800e+00"
    assert "%.2e" % r["pri_res"] == "1.91e-07"
    assert "%.2e" % r["dua_res"] == "7.50e-07"
    assert "%.2e" % r["rho"] == "1.38e+00"
    np.testing.assert_allclose(r["x"], [0.3, 0.7], atol=1e-6)
    np.testing.assert_allclose(r["y"], [-2.9, 0.0, 0.2], atol=1e-6)


@pytest.mark.parametrize("linsys", [0, 1])
def test_osqp_demo_default_alpha(oracle, linsys):
    r = _demo(oracle, default_params(), max_iter=4000, linsys=linsys)
    assert r["status"] == 1 and r["iters"] == 25
    np.testing.assert_allclose(r["x"], [0.3, 0.7], atol=5e-3)


def test_osqp_primal_infeasible_and_maxiter(oracle):
    p = default_params()
    # x >= 1 and x <= 0 at once
    r = oracle.osqp_solve(p, 1, 2, [0, 1], [0], [1.0], [0.0], [0, 2], [0, 1], [1.0, 1.0],
                          [1.0, -1e30], [1e30, 0.0], max_iter=4000)
    assert r["status"] == -3 and np.all(np.isnan(r["x"]))
    r = _demo(oracle, p, max_iter=3)
    assert r["status"] == -2 and r["iters"] == 3


def _random_qp(rng, n, m):
    M = rng.standard_normal((n, n))
    P = M @ M.T + 0.1 * np.eye(n)
    A = rng.standard_normal((m, n)) * (rng.random((m, n)) < 0.5)
    x0 = rng.standard_normal(n)
    l = A @ x0 - rng.random(m)
    u = A @ x0 + rng.random(m)
    eq = rng.random(m) < 0.2
    l[eq] = u[eq] = (A @ x0)[eq]
    q = rng.standard_normal(n)
    return P, q, A, l, u


def _csc(M, upper=False):
    n_col = M.shape[1]
    p, i, x = [0], [], []
    for j in range(n_col):
        for r in range(M.shape[0]):
            if M[r, j] != 0 and (not upper or r <= j):
                i.append(r); x.append(M[r, j])
        p.append(len(i))
    return np.asarray(p, np.int32), np.asarray(i, np.int32), np.asarray(x, np.float64)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_osqp_random_qp_kkt(oracle, seed):
    """Converged solutions satisfy the QP's optimality conditions (numpy check)."""
    rng = np.random.default_rng(seed)
    n, m = 12, 20
    P, q, A, l, u = _random_qp(rng, n, m)
    p = default_params()
    p.eps_abs = p.eps_rel = 1e-7
    Pp, Pi, Px = _csc(P, upper=True)
    Ap, Ai, Ax = _csc(A)
    out = {}
    for linsys in (0, 1):
        r = oracle.osqp_solve(p, n, m, Pp, Pi, Px, q, Ap, Ai, Ax, l, u, max_iter=20000, linsys=linsys)
        assert r["status"] == 1
        x, y = r["x"], r["y"]
        assert np.abs(P @ x + q + A.T @ y).max() < 1e-5          # stationarity
        Axv = A @ x
        assert np.all(Axv >= l - 1e-5) and np.all(Axv <= u + 1e-5)  # primal feasibility
        act_lo, act_hi = np.abs(Axv - l) < 1e-4, np.abs(Axv - u) < 1e-4
        assert np.all(np.abs(y[~act_lo & ~act_hi]) < 1e-4)       # complementarity
        assert np.all(y[act_lo & ~act_hi] <= 1e-6) and np.all(y[act_hi & ~act_lo] >= -1e-6)
        out[linsys] = x
    np.testing.assert_allclose(out[0], out[1], atol=1e-8)


def _agent_qp(oracle, p, ins, a, corr=None):
    g = ins.guess[a]
    nt = g.shape[1]
    if corr is None:
        corr, _, _ = oracle.agent_corridors(p, g[0], g[1], g[2], ins.dimx, ins.dimy, ins.obstacles, False)
    cfg = np.array([g[0, 0], g[0, -1], g[1, 0], g[1, -1], g[2, 0], g[2, -1]])
    qp = oracle.assemble_qp(p, g, g[:2], cfg, corr, ins.plane_t[a], ins.plane_abc[a])
    return qp, corr, nt


def _dense(qp):
    A = np.zeros((qp["m"], qp["n"]))
    for j in range(qp["n"]):
        for k in range(qp["Ap"][j], qp["Ap"][j + 1]):
            A[qp["Ai"][k], j] += qp["Ax"][k]
    P = np.zeros((qp["n"], qp["n"]))
    for j in range(qp["n"]):
        for k in range(qp["Pp"][j], qp["Pp"][j + 1]):
            P[qp["Pi"][k], j] = qp["Px"][k]
            P[j, qp["Pi"][k]] = qp["Px"][k]
    return A, P


def test_qp_assembly_invariants(oracle, params):
    """Sizes and the reference's own diagnostics as assertions."""
    p = params
    ins = synthetic_instance(5, 50.0, 6, 20, (8, 14), p)
    ins.plane_t, ins.plane_abc, _ = oracle.instance_planes(p, ins.guess)
    for a in range(ins.n_agents):
        qp, corr, nt = _agent_qp(oracle, p, ins, a)
        K = len(ins.plane_t[a])
        assert qp["n"] == 6 * nt - 2 and qp["m"] == 13 * nt + 4 * K
        assert qp["Ap"][-1] == 28 * nt - 11 + 12 * K
        A, P = _dense(qp)
        g = ins.guess[a]
        s0 = np.concatenate([g[0], g[1], g[2], g[3], g[4, :-1], g[5, :-1]])
        r = A @ s0
        nm, dt = nt - 1, p.dt
        # (i) kinematic rows at the guess == nonlinear defect (dsqp_solver.cc:726-729)
        x, y, yaw, st, v, w = g[0], g[1], g[2], g[3], g[4, :-1], g[5, :-1]
        fx = x[:-1] + dt * v * np.cos(yaw[:-1]) - x[1:]
        fy = y[:-1] + dt * v * np.sin(yaw[:-1]) - y[1:]
        fp = yaw[:-1] + dt * v * np.tan(st[:-1]) / p.WB - yaw[1:]
        fs = st[:-1] + dt * w - st[1:]
        kin = r[:4 * nm] - qp["l"][:4 * nm]
        np.testing.assert_allclose(kin, np.concatenate([fx, fy, fp, fs]), atol=1e-9)
        assert np.array_equal(qp["l"][:4 * nm + 6], qp["u"][:4 * nm + 6])
        # (ii) cfg rows exactly satisfied (:775-782)
        si = 4 * nm
        assert np.all(r[si:si + 6] == qp["l"][si:si + 6])
        # (iii) corridor slack >= 0: every box grows around its own (float) centre (:954-963)
        si += 6
        assert np.all(r[si:si + 4 * nt] - qp["l"][si:si + 4 * nt] > -1e-5)
        assert np.all(qp["u"][si:si + 4 * nt] - r[si:si + 4 * nt] > -1e-5)
        # (iv) trust slack == r_trust at SQP iteration 0 (:991-992)
        si += 4 * nt
        np.testing.assert_allclose(r[si:si + 2 * nt] - qp["l"][si:si + 2 * nt], p.r_trust, atol=1e-12)
        # ctrl / steer bounds
        si += 2 * nt
        assert np.all(qp["u"][si:si + nm] == p.max_v) and np.all(qp["l"][si + nm:si + 2 * nm] == -p.max_omega)
        assert np.all(qp["u"][si + 2 * nm:si + 2 * nm + nt] == p.steer_max)
        # (v) inter rows: l = -inf, slack >= 0 at a collision-free guess (:1121-1128)
        si += 2 * nm + nt
        assert np.all(np.isneginf(qp["l"][si:]))
        assert np.all(qp["u"][si:] - r[si:] > -1e-9)
        # objective = 1/2 sum dv^2 + 1/2 sum w^2 (:163-197)
        z = np.random.default_rng(a).standard_normal(qp["n"])
        vv, ww = z[4 * nt:4 * nt + nm], z[4 * nt + nm:]
        np.testing.assert_allclose(0.5 * z @ P @ z, 0.5 * np.sum(np.diff(vv) ** 2) + 0.5 * np.sum(ww ** 2))
        # time-major ordering makes P + A' A banded with half-bandwidth 6
        perm = []
        for t in range(nt):
            perm += [t, nt + t, 2 * nt + t, 3 * nt + t] + ([4 * nt + t, 4 * nt + nm + t] if t < nm else [])
        Af = np.where(np.isfinite(A), A, 0)
        H = (P + Af.T @ Af)[np.ix_(perm, perm)]
        ii, jj = np.nonzero(H)
        assert np.abs(ii - jj).max() == 6


def test_corridor_boxes_golden_and_properties(oracle, params):
    gd = np.load(os.path.join(GOLD, "corridor_boxes_golden.npz"))
    for (x, y), box, st in zip(gd["pts"], gd["boxes"], gd["status"]):
        b, s = oracle.generate_box(params, 50.0, 50.0, float(x), float(y), gd["obs"])
        assert np.array_equal(b, box) and np.array_equal(s, st)
    # free space: grows to the accumulated-0.1 limit; 100 additions of 0.1 stay below 10 (App. C)
    b, s = oracle.generate_box(params, 100.0, 100.0, 50.0, 50.0, np.zeros((0, 3)))
    acc = 0.0
    for _ in range(101):
        acc += 0.1
    assert acc > 10.0 and list(s) == [1, 0]
    np.testing.assert_allclose(b, [50 - 10.1, 50 - 10.1, 50 + 10.1, 50 + 10.1], atol=1e-9)
    # outside the map -> projected, initial_status 1
    b, s = oracle.generate_box(params, 50.0, 50.0, 0.5, 25.0, np.zeros((0, 3)))
    assert s[1] == 1 and b[0] >= params.rv
    # inside an obstacle's inflated square -> legal point search, initial_status 2; box avoids it
    obs = np.array([[25.0, 25.0, 0.8]])
    b, s = oracle.generate_box(params, 50.0, 50.0, 25.5, 25.2, obs)
    R = 0.8 + params.rv
    assert s[1] == 2 and s[0] == 1
    assert not (b[0] - R < 25.0 < b[2] + R and b[1] - R < 25.0 < b[3] + R)


def test_planes_properties(oracle, params):
    ins = synthetic_instance(9, 50.0, 8, 10, (8, 14), params)
    pts, pabc, legal = oracle.instance_planes(params, ins.guess)
    assert legal
    f32 = np.float32
    total = 0
    for a in range(ins.n_agents):
        t = pts[a]
        assert np.all(np.diff(t) >= 0)   # sorted by time (push order)
        total += len(t)
        g = ins.guess[a]
        for k in range(len(t)):
            tt = t[k]
            xf = float(f32(g[0, tt] + params.f2x * np.cos(g[2, tt])))
            yf = float(f32(g[1, tt] + params.f2x * np.sin(g[2, tt])))
            a_, b_, c_ = pabc[a][k, 0:3]
            # own front disc is on the negative side with margin (rv offset bisector)
            assert a_ * xf + b_ * yf + c_ <= 1e-9
    assert total % 2 == 0 and total > 0


def test_interpolation_matches_oracle(oracle, params):
    from csdotrajectoryplanning_b200.scenario import interpolate_initial_guess, _primitive
    rng = np.random.default_rng(3)
    for trial in range(5):
        s = np.array([20.0, 20.0, rng.uniform(-3, 3)])
        states, acts = [s], []
        for k in range(10):
            a = int(rng.integers(0, 3)) if k != 4 else 6
            states.append(states[-1].copy() if a == 6 else _primitive(states[-1], a))
            acts.append(a)
        states = np.asarray(states)
        goal = states[-1] + np.array([0.01, -0.02, 0.005])
        g = interpolate_initial_guess([(states, acts)], goal[None, :], params)[0]
        go, ns = oracle.interpolate_guess(states, acts, goal, 2, params.dt, 3.0, params.LF, params.LB, g.shape[1] + 4)
        assert ns == g.shape[1] == 31
        np.testing.assert_allclose(go[:, :ns], g, atol=1e-12)
        assert np.all(go[3:, ns:] == 0) and np.all(go[0, ns:] == go[0, ns - 1])


def _load_golden():
    from csdotrajectoryplanning_b200.batch import Batch
    gd = np.load(os.path.join(GOLD, "dsqp_refine_golden.npz"))
    b = Batch(*(gd[k] for k in ("inst_agent_ptr", "inst_nt", "inst_dims", "obs_ptr", "obs", "agent_off",
                                "guess", "plane_ptr", "plane_t", "plane_abc")))
    return b, gd


@pytest.mark.parametrize("linsys", [0, 1])
def test_refine_golden(oracle, params, linsys):
    b, gd = _load_golden()
    res, _ = oracle.refine(params, b, linsys=linsys, nthreads=2)
    for k in ("status", "sqp_iters", "admm_iters", "n_factor", "inst_status", "inst_static_legal"):
        assert np.array_equal(getattr(res, k), gd[k]), k
    # tolerance of north_star: 1e-3 m / 1e-3 rad (observed: ~1e-8 between the two linear solvers)
    assert np.abs(res.traj - gd["traj"]).max() < 1e-6
    assert np.abs(res.corridors - gd["corridors"]).max() < 1e-6
    assert np.abs(res.objective - gd["objective"]).max() < 1e-8


def test_refine_results_are_sane(oracle, params, small_batch):
    res, flops = oracle.refine(params, small_batch, linsys=1, nthreads=2)
    assert flops > 0 and np.all(np.abs(res.status) <= 2)
    for a in range(small_batch.n_agents):
        g, tr = small_batch.agent_guess(a), res.agent_traj(small_batch, a)
        # start/goal pinned (cfg rows), trust region respected, actuator limits respected
        assert np.abs(tr[:3, 0] - g[:3, 0]).max() < 5e-3 and np.abs(tr[:3, -1] - g[:3, -1]).max() < 5e-3
        # OSQP stops at eps_rel = 1e-3 relative to |z| ~ 50 m, so bounds hold to ~0.1 m only
        assert np.abs(tr[:2] - g[:2]).max() < params.r_trust + 0.15
        assert np.abs(tr[4]).max() < params.max_v + 5e-2 and np.abs(tr[5]).max() < params.max_omega + 5e-2


def test_status_aggregation_rule(oracle, params):
    """dsqp_solver.cc:1224-1238 evaluated literally: once a |s|>2 code is recorded the last one wins."""
    def agg(sts):
        worst, anyu = 2, False
        for s in sts:
            if abs(s) > 1:
                anyu = True
                if abs(s) > worst:
                    worst = s
        return worst if anyu else 1
    assert agg([1, 1]) == 1 and agg([1, 2, 1]) == 2 and agg([-2, 1]) == 2
    assert agg([-3, 1, 2]) == 2 and agg([1, -3]) == -3 and agg([3, -2, 1]) == 3
    assert agg([-3, -2]) == -2  # a negative running value lets any later record replace it
