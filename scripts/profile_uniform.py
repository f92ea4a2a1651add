"""Developer tooling: uniform-horizon workload (all agents Nt = 3*acts+1) for phase profiling."""
import sys
sys.path.insert(0, ".")
from csdotrajectoryplanning_b200 import default_params, pack_instances
from csdotrajectoryplanning_b200.scenario import synthetic_batch
from csdotrajectoryplanning_b200.solver import DsqpSolver
acts = int(sys.argv[1]) if len(sys.argv) > 1 else 30
n = int(sys.argv[2]) if len(sys.argv) > 2 else 24
na = int(sys.argv[3]) if len(sys.argv) > 3 else 25
p = default_params()
inst = synthetic_batch([(50.0, na, 25, (acts, acts))], n, seed=77, params=p)
S = DsqpSolver(p)
b, _ = S.planes(pack_instances(inst))
r = S.refine(b)
import numpy as np
print("K mean", float(np.diff(b.plane_ptr).mean()), "agents", b.n_agents, "Nt", set(b.inst_nt.tolist()), "admm", int(r.admm_iters.sum()), "nfac", int(r.n_factor.sum()), "qps", int(r.n_qp.sum()))
