"""Developer tooling: a tiny refine (with planes, obstacles, ragged horizons) for compute-sanitizer runs.
usage: sanitize_small.py [short|long]   short: horizons 25..88 (one-warp solver variants), long: horizon ~127 and
~190 (CTA-wide solver variants, 128- and 192-thread CTAs)"""
import sys
sys.path.insert(0, ".")
from csdotrajectoryplanning_b200 import default_params, pack_instances
from csdotrajectoryplanning_b200.scenario import synthetic_instance
from csdotrajectoryplanning_b200.solver import DsqpSolver
p = default_params()
which = sys.argv[1] if len(sys.argv) > 1 else "short"
if which == "short":
    inst = [synthetic_instance(7, 50.0, 4, 10, (8, 12), p), synthetic_instance(8, 50.0, 3, 0, (27, 29), p)]
elif which == "dense":   # a small map: many inter-vehicle planes per agent (plane-major passes)
    inst = [synthetic_instance(11, 25.0, 6, 4, (9, 12), p)]
elif which == "mid":     # horizon ~140: stride 144 in 160-thread CTAs (stride != block size), two CTAs per SM
    inst = [synthetic_instance(12, 70.0, 3, 6, (46, 47), p)]
elif which == "long":
    inst = [synthetic_instance(9, 60.0, 3, 6, (41, 42), p)]
else:
    inst = [synthetic_instance(10, 80.0, 2, 6, (62, 63), p)]
S = DsqpSolver(p)
b, _ = S.planes(pack_instances(inst))
r = S.refine(b)
print(which, "status", r.status.tolist(), "sqp", r.sqp_iters.tolist(), "admm", int(r.admm_iters.sum()), "planes", int(b.plane_ptr[-1]), S.last_launch())
