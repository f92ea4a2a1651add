"""GPU test at BASELINE configs[1] size (600 instances / 9000 agents of the map50by50 sweep shape, the bench
workload): size-independent properties instead of an oracle comparison -- bit determinism, instance-order
invariance, the status aggregation rule, the QP's own constraints on every solved trajectory, and the
iteration totals that every parity-green build so far has produced (the CPU oracle gives the same
totals on this workload: bench.py's cpu_baseline leg)."""
import copy

import numpy as np
import pytest

from csdotrajectoryplanning_b200 import pack_instances
from csdotrajectoryplanning_b200.scenario import MAP50_SWEEP, synthetic_batch
from csdotrajectoryplanning_b200.sharding import aggregate_instance_status

pytestmark = pytest.mark.gpu


def _check_solved_agents_respect_their_qp(b, r1, params, extent):
    """Every agent whose last QP was solved (status 1) satisfies that QP's own box constraints to OSQP's primal
    tolerance eps_abs + eps_rel * max(|Ax|, |z|)_inf on ANY row: positions reach `extent` m and the inter-vehicle
    rows are not normalised (|a|, |b| up to 2 sqrt(2) r_trust), so the reference's own criterion admits
    violations of several 1e-2 on agents with planes.  Returns the number of agents checked."""
    n_checked = 0
    for a in range(b.n_agents):
        if int(r1.status[a]) != 1:
            continue
        k0, k1 = int(b.plane_ptr[a]), int(b.plane_ptr[a + 1])
        rowmax = extent
        if k1 > k0:
            abc = b.plane_abc[12 * k0:12 * k1].reshape(-1, 3)
            rowmax = max(rowmax, float((np.abs(abc[:, 0]) + np.abs(abc[:, 1])).max()) * (extent + 2.0))
        tol = 1.2e-3 * (1.0 + rowmax) + 5e-3
        o, nt = int(b.agent_off[a]), int(b.agent_off[a + 1] - b.agent_off[a])
        g = b.guess[6 * o:6 * (o + nt)].reshape(6, nt)
        x = r1.agent_traj(b, a)
        assert np.abs(x[:3, 0] - g[:3, 0]).max() < tol and np.abs(x[:3, -1] - g[:3, -1]).max() < tol    # cfg rows
        assert np.abs(x[:2] - g[:2]).max() <= params.r_trust + tol                                       # trust region
        assert np.abs(x[3]).max() <= params.steer_max + tol, (a, np.abs(x[3]).max())
        assert np.abs(x[4, :-1]).max() <= params.max_v + tol, (a, np.abs(x[4, :-1]).max())
        assert np.abs(x[5, :-1]).max() <= params.max_omega + tol, (a, np.abs(x[5, :-1]).max())
        n_checked += 1
    return n_checked


def test_full_sweep_properties(params, solver):
    inst = synthetic_batch(MAP50_SWEEP, 60, seed=1234, params=params)
    b, _ = solver.planes(pack_instances(inst))
    assert b.n_inst == 600 and b.n_agents == 9000
    r1 = solver.refine(b)
    r2 = solver.refine(b)
    for k in ("traj", "corridors", "status", "sqp_iters", "n_qp", "admm_iters", "n_factor", "objective",
              "inst_status", "inst_static_legal"):
        assert np.array_equal(getattr(r1, k), getattr(r2, k)), k          # bit-deterministic
    # totals of the bench workload
    assert int(r1.n_qp.sum()) == 40284 and int(r1.admm_iters.sum()) == 4976900 and int(r1.n_factor.sum()) == 76201
    # aggregation rule of dsqp_solver.cc:1224-1243 on the host from the per-agent statuses
    chk = copy.deepcopy(r1)
    aggregate_instance_status(b, chk)
    assert np.array_equal(chk.inst_status, r1.inst_status)
    # bookkeeping
    assert np.array_equal(r1.n_qp, r1.sqp_iters) and r1.sqp_iters.max() <= params.max_iter and r1.sqp_iters.min() >= 1
    assert np.all(r1.admm_iters <= r1.n_qp * params.osqp_max_iter) and np.all(r1.n_factor >= r1.n_qp)
    n_checked = _check_solved_agents_respect_their_qp(b, r1, params, 52.0)
    assert n_checked > 8000
    # instance-order invariance: the first 40 instances, reversed, give the same per-instance results
    sub = list(reversed(inst[:40]))
    bs, _ = solver.planes(pack_instances(sub))
    rs = solver.refine(bs)
    for j, ins in enumerate(sub):
        i = 39 - j
        assert rs.inst_status[j] == r1.inst_status[i]
        a0, a1 = int(b.inst_agent_ptr[i]), int(b.inst_agent_ptr[i + 1])
        s0 = int(bs.inst_agent_ptr[j])
        for a in range(a0, a1):
            assert np.array_equal(r1.agent_traj(b, a), rs.agent_traj(bs, s0 + a - a0))


def test_c5_shape_properties(params, solver):
    """BASELINE configs[4] shape (100x100 maps, 100 agents, 50 obstacles, horizons 127 / 190 / 256; 24 instances,
    2400 agents, ~200 planes per agent): determinism, status rule, bookkeeping, the QP's own constraints, and
    independence of the launch shape -- the horizon-127 instances refined on their own run in 128-thread CTAs
    (2 per SM), inside the mixed batch in 256-thread CTAs (1 per SM), with the same result."""
    from tools import synth
    inst = synth.synth_batch(synth.C5_SHAPES, 8, 1234, params)
    b, _ = solver.planes(pack_instances(inst))
    assert b.n_inst == 24 and b.n_agents == 2400 and sorted(set(b.inst_nt.tolist())) == [127, 190, 256]
    assert b.plane_ptr[-1] > 100 * b.n_agents
    r1 = solver.refine(b)
    assert solver.last_launch()["block"] == 256
    r2 = solver.refine(b)
    for k in ("traj", "corridors", "status", "sqp_iters", "n_qp", "admm_iters", "n_factor", "objective",
              "inst_status", "inst_static_legal"):
        assert np.array_equal(getattr(r1, k), getattr(r2, k)), k
    chk = copy.deepcopy(r1)
    aggregate_instance_status(b, chk)
    assert np.array_equal(chk.inst_status, r1.inst_status)
    assert np.array_equal(r1.n_qp, r1.sqp_iters) and r1.sqp_iters.max() <= params.max_iter and r1.sqp_iters.min() >= 1
    assert np.all(r1.admm_iters <= r1.n_qp * params.osqp_max_iter) and np.all(r1.n_factor >= r1.n_qp)
    assert _check_solved_agents_respect_their_qp(b, r1, params, 102.0) > 1500
    short = [i for i in range(b.n_inst) if b.inst_nt[i] == 127]
    bs, _ = solver.planes(pack_instances([inst[i] for i in short]))
    rs = solver.refine(bs)
    assert solver.last_launch()["block"] == 128 and solver.last_launch()["ctas_per_sm"] == 2
    worst = 0.0
    for j, i in enumerate(short):
        a0, a1, s0 = int(b.inst_agent_ptr[i]), int(b.inst_agent_ptr[i + 1]), int(bs.inst_agent_ptr[j])
        assert rs.inst_status[j] == r1.inst_status[i]
        for a in range(a0, a1):
            for k in ("status", "sqp_iters", "admm_iters", "n_factor"):
                assert getattr(r1, k)[a] == getattr(rs, k)[s0 + a - a0], (k, a)
            worst = max(worst, float(np.abs(r1.agent_traj(b, a) - rs.agent_traj(bs, s0 + a - a0)).max()))
    assert worst < 1e-7, worst
