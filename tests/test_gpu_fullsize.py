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
    # every agent whose last QP was solved (status 1) satisfies that QP's own box constraints
    # to OSQP's primal tolerance eps_abs + eps_rel * max(|Ax|, |z|)_inf on ANY row: positions reach ~52 m
    # and the inter-vehicle rows are not normalised (|a|, |b| up to 2 sqrt(2) r_trust), so the reference's
    # own criterion admits violations of several 1e-2 on agents with planes
    n_checked = 0
    for a in range(b.n_agents):
        if int(r1.status[a]) != 1:
            continue
        k0, k1 = int(b.plane_ptr[a]), int(b.plane_ptr[a + 1])
        rowmax = 52.0
        if k1 > k0:
            abc = b.plane_abc[12 * k0:12 * k1].reshape(-1, 3)
            rowmax = max(rowmax, float((np.abs(abc[:, 0]) + np.abs(abc[:, 1])).max()) * 54.0)
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
