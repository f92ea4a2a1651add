"""Developer tooling: a tiny refine (with planes, obstacles, ragged horizons) for compute-sanitizer runs."""
import sys
sys.path.insert(0, ".")
from csdotrajectoryplanning_b200 import default_params, pack_instances
from csdotrajectoryplanning_b200.scenario import synthetic_instance
from csdotrajectoryplanning_b200.solver import DsqpSolver
p = default_params()
inst = [synthetic_instance(7, 50.0, 4, 10, (8, 12), p), synthetic_instance(8, 50.0, 3, 0, (27, 29), p)]
S = DsqpSolver(p)
b, _ = S.planes(pack_instances(inst))
r = S.refine(b)
print("status", r.status.tolist(), "sqp", r.sqp_iters.tolist(), "admm", int(r.admm_iters.sum()), S.last_launch())
