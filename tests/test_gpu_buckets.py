"""Horizon buckets (csdo_refine, csdo_refine_device_hinted): in a batch of mixed horizons every horizon class gets
its own launch.  Checked against the single launch shaped for the longest horizon, against the oracle, and
between the host-buffer and the device-resident entry points."""
import os

import numpy as np
import pytest

from csdotrajectoryplanning_b200 import pack_instances
from csdotrajectoryplanning_b200.scenario import synthetic_instance
from csdotrajectoryplanning_b200.solver import DeviceBatch, DeviceResult

pytestmark = pytest.mark.gpu

COUNTERS = ("status", "sqp_iters", "n_qp", "admm_iters", "n_factor", "inst_status", "inst_static_legal")


def _mixed_batch(params, oracle=None):
    """Horizons 28..190 in one batch: classes 64, 96, 128, 192 (both solver families)."""
    spec = [(600, 40.0, 6, 8, (9, 10)), (601, 40.0, 6, 8, (17, 19)), (602, 50.0, 6, 6, (27, 30)),
            (603, 60.0, 5, 6, (40, 42)), (604, 80.0, 4, 6, (61, 63)), (605, 40.0, 6, 0, (12, 14))]
    inst = [synthetic_instance(s, size, na, no, acts, params) for s, size, na, no, acts in spec]
    if oracle is not None:
        for ins in inst:
            ins.plane_t, ins.plane_abc, _ = oracle.instance_planes(params, ins.guess)
    return inst


@pytest.fixture(autouse=True)
def _every_class_its_own_launch():
    """The test batches are tiny: without this, classes below ~4 agents per SM are merged upwards."""
    os.environ["CSDO_BUCKET_MIN"] = "1"
    yield
    os.environ.pop("CSDO_BUCKET_MIN", None)


def _per_agent_diff(b, r0, r1):
    return np.array([np.abs(r0.agent_traj(b, a) - r1.agent_traj(b, a)).max() for a in range(b.n_agents)])


def test_buckets_equal_single_launch_and_oracle(oracle, params, solver):
    inst = _mixed_batch(params, oracle)
    b = pack_instances(inst)
    nt = b.agent_nt()
    assert nt.min() < 64 and nt.max() > 160 and b.plane_ptr[-1] > 0
    rb = solver.refine(b)
    info = solver.last_launch()
    assert info["block"] >= nt.max() and info["launches"] == 2 + 2 * 4     # classes 64, 96, 128, 192
    os.environ["CSDO_BUCKET_MIN"] = "100000"       # merging: one launch per solver family, same bits
    rm = solver.refine(b)
    assert solver.last_launch()["launches"] == 2 + 2 * 2
    os.environ["CSDO_BUCKET_MIN"] = "1"
    for k in COUNTERS + ("traj", "corridors"):
        assert np.array_equal(getattr(rm, k), getattr(rb, k)), k
    os.environ["CSDO_NO_BUCKETS"] = "1"
    try:
        r1 = solver.refine(b)
        assert solver.last_launch()["launches"] == 4
    finally:
        del os.environ["CSDO_NO_BUCKETS"]
    for k in COUNTERS:
        assert np.array_equal(getattr(rb, k), getattr(r1, k)), k
    d = _per_agent_diff(b, rb, r1)
    # above 96 steps both runs use the CTA-wide solver, whose arithmetic does not depend on the launch shape
    assert np.all(d[nt > 96] == 0.0)
    # up to 96 steps the bucket runs the one-warp solver, the single launch the CTA-wide one: rounding only
    assert d.max() < 1e-6
    ro, _ = oracle.refine(params, b, linsys=0, nthreads=4)
    for k in COUNTERS:
        assert np.array_equal(getattr(ro, k), getattr(rb, k)), k
    assert np.abs(ro.traj - rb.traj).max() < 1e-5


def test_device_hinted_equals_host_path_and_supports_subsets(params, solver):
    import torch
    inst = _mixed_batch(params)
    b, _ = solver.planes(pack_instances(inst))
    rh = solver.refine(b)                                   # host buffers: bucketed inside csdo_refine
    dev = torch.device("cuda", 0)
    db, dr = DeviceBatch(b, dev), DeviceResult(b, dev)
    solver.refine_device(db, dr); solver.sync()             # csdo_refine_device_hinted
    assert solver.last_launch()["launches"] >= 8
    rd = dr.to_host()
    for k in COUNTERS + ("traj", "corridors", "objective"):
        assert np.array_equal(getattr(rd, k), getattr(rh, k)), k
    dr1 = DeviceResult(b, dev)
    solver.refine_device(db, dr1, by_horizon=False); solver.sync()   # one launch for the longest horizon
    assert solver.last_launch()["launches"] == 4
    r1 = dr1.to_host()
    for k in COUNTERS:
        assert np.array_equal(getattr(r1, k), getattr(rh, k)), k
    assert _per_agent_diff(b, r1, rh).max() < 1e-6
    # a subset of the agents (the agent-partitioned mode): those agents get the same bits, the others are untouched
    ids = np.arange(b.n_agents, dtype=np.int32)[1::2]
    db2, dr2 = DeviceBatch(b, dev), DeviceResult(b, dev)
    db2.set_active(ids)
    solver.refine_device(db2, dr2); solver.sync()
    r2 = dr2.to_host()
    for a in range(b.n_agents):
        if a in set(ids.tolist()):
            assert np.array_equal(r2.agent_traj(b, a), rh.agent_traj(b, a)) and r2.admm_iters[a] == rh.admm_iters[a]
        else:
            assert not r2.agent_traj(b, a).any() and r2.sqp_iters[a] == 0
