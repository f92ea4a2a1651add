"""dev: bucketed vs single-launch refine on a mixed batch: result differences."""
import os, sys
sys.path.insert(0, ".")
import numpy as np
import bench
from csdotrajectoryplanning_b200 import default_params, pack_instances
from csdotrajectoryplanning_b200.solver import DsqpSolver
from tools import synth
p = default_params()
S = DsqpSolver(p)
for name, inst in (("real", bench.build_instances("real", 579, 0, 1, p)), ("c5", synth.synth_batch(synth.C5_SHAPES, 8, 1234, p))):
    b, _ = S.planes(pack_instances(inst))
    r1 = S.refine(b); l1 = S.last_launch()
    os.environ["CSDO_NO_BUCKETS"] = "1"
    r0 = S.refine(b); l0 = S.last_launch()
    del os.environ["CSDO_NO_BUCKETS"]
    d = np.array([np.abs(r0.agent_traj(b, a) - r1.agent_traj(b, a)).max() for a in range(b.n_agents)])
    nt = b.agent_nt()
    print(name, "agents", b.n_agents, "counters equal", all(np.array_equal(getattr(r0, k), getattr(r1, k)) for k in ("status", "sqp_iters", "admm_iters", "n_factor", "inst_status")),
          "bit-equal agents", int((d == 0).sum()), "max", d.max(), "max over Nt>96", d[nt > 96].max() if (nt > 96).any() else None,
          "bit-equal among Nt>96", int((d[nt > 96] == 0).sum()), "of", int((nt > 96).sum()), l1, l0)
