"""Small fixed workload for ncu captures (developer tooling): N instances of the map50by50 sweep shape."""
import sys
import numpy as np
sys.path.insert(0, ".")
from csdotrajectoryplanning_b200 import default_params, pack_instances
from csdotrajectoryplanning_b200.scenario import MAP50_SWEEP, synthetic_batch
from csdotrajectoryplanning_b200.solver import DsqpSolver

per_shape = int(sys.argv[1]) if len(sys.argv) > 1 else 4
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
p = default_params()
inst = synthetic_batch(MAP50_SWEEP, per_shape, seed=1234, params=p)
S = DsqpSolver(p)
b, _ = S.planes(pack_instances(inst))
import time
for _ in range(reps):
    t = time.time(); r = S.refine(b); dt = time.time() - t
print("agents", b.n_agents, "qps", int(r.n_qp.sum()), "admm", int(r.admm_iters.sum()), "nfac", int(r.n_factor.sum()),
      "time %.3f s" % dt, S.last_launch())
