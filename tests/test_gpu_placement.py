"""Placement / variant knobs of the refine launch against the default launch on the same batch: where the row
arrays, the band factor and the plane records live (Layout.tier, CSDO_EXTRA_SMEM) and which register class runs
(CSDO_NO_LEAN) must not change the result."""
import os

import numpy as np
import pytest

from csdotrajectoryplanning_b200 import pack_instances
from csdotrajectoryplanning_b200.scenario import synthetic_instance

pytestmark = pytest.mark.gpu

CASES = {"one_warp_solver_nt_le_96": [(520, 40.0, 8, 10, (9, 15)), (521, 40.0, 7, 0, (24, 29))],
         "cta_wide_solver_nt_127": [(530, 60.0, 6, 8, (41, 42))]}


@pytest.mark.parametrize("env", [{"CSDO_TIER": "0"}, {"CSDO_TIER": "3"}, {"CSDO_EXTRA_SMEM": "1"}, {"CSDO_NO_LEAN": "1"},
                                 {"CSDO_MAX_CTAS_PER_SM": "1"}], ids=lambda e: "-".join(f"{k}={v}" for k, v in e.items()))
@pytest.mark.parametrize("case", sorted(CASES))
def test_placement_variants_agree(params, solver, env, case):
    inst = [synthetic_instance(s, size, na, no, acts, params) for s, size, na, no, acts in CASES[case]]
    b, _ = solver.planes(pack_instances(inst))
    assert b.plane_ptr[-1] > 0
    ref = solver.refine(b)
    base = solver.last_launch()
    os.environ.update(env)
    try:
        got = solver.refine(b)
        info = solver.last_launch()
    finally:
        for k in env:
            del os.environ[k]
    for k in ("status", "sqp_iters", "n_qp", "admm_iters", "n_factor", "inst_status", "inst_static_legal"):
        assert np.array_equal(getattr(got, k), getattr(ref, k)), k
    # same arithmetic, different homes for the data (and, for the lean / non-lean pair, another register allocation
    # of the same expressions): the iterates agree to rounding
    assert np.abs(got.traj - ref.traj).max() < 1e-7 and np.abs(got.corridors - ref.corridors).max() < 1e-7
    if "CSDO_TIER" in env:
        assert info["tier"] == int(env["CSDO_TIER"])
    if "CSDO_EXTRA_SMEM" in env:
        assert info["smem_bytes"] >= base["smem_bytes"]
    if "CSDO_MAX_CTAS_PER_SM" in env:
        assert info["ctas_per_sm"] == 1
