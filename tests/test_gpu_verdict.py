"""GPU test of north_star's verdict criteria: planning success flag and collision verdict identical
between the CUDA path and the CPU oracle on the same inputs (after dumpSolutions' 3-decimal rounding)."""
import numpy as np
import pytest

from conftest import make_batch
from csdotrajectoryplanning_b200 import verdict as V

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seeds,na,no", [([201, 202, 203, 204], 6, 20), ([211, 212], 12, 25)])
def test_success_and_collision_verdict_identical(oracle, params, solver, seeds, na, no):
    b = make_batch(oracle, params, seeds, na=na, no=no, acts=(10, 20))
    ro, _ = oracle.refine(params, b, linsys=0, nthreads=4)
    rg = solver.refine(b)
    for i in range(b.n_inst):
        assert V.success(ro.inst_status[i]) == V.success(rg.inst_status[i])
        a0, a1 = int(b.inst_agent_ptr[i]), int(b.inst_agent_ptr[i + 1])
        obs = b.obs[3 * b.obs_ptr[i]:3 * b.obs_ptr[i + 1]].reshape(-1, 3)
        to = [V.rounded_solution(ro.agent_traj(b, a)) for a in range(a0, a1)]
        tg = [V.rounded_solution(rg.agent_traj(b, a)) for a in range(a0, a1)]
        assert V.verdict(to, obs) == V.verdict(tg, obs)
        # the printed (3-decimal) trajectories themselves agree except where a value sits on a rounding edge
        diff = max(np.abs(x - y).max() for x, y in zip(to, tg))
        assert diff <= 1e-3 + 1e-12
