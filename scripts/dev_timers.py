"""Developer tooling: per-phase cycle timers of one workload shape (needs a -DCSDO_DEV_TIMERS library, e.g.
bash scripts/dev_build.sh timers -DCSDO_DEV_TIMERS; CSDO_LIB=.../csrc/_dev/libcsdo_timers.so CSDO_PROFILE=1).
usage: dev_timers.py <c5|c5a|c5b|c5c|map50> <instances per shape>
The library prints cycles summed over CTAs (thread 0's clock); this prints the ADMM iteration count to divide by."""
import os, sys
sys.path.insert(0, ".")
from csdotrajectoryplanning_b200 import default_params, pack_instances
from csdotrajectoryplanning_b200.solver import DsqpSolver
from tools import synth

name, per = sys.argv[1], int(sys.argv[2])
p = default_params()
if name.startswith("c5"):
    shapes = {"c5": synth.C5_SHAPES, "c5a": synth.C5_SHAPES[:1], "c5b": synth.C5_SHAPES[1:2], "c5c": synth.C5_SHAPES[2:]}[name]
    inst = synth.synth_batch(shapes, per, 1234, p)
else:
    from csdotrajectoryplanning_b200.scenario import MAP50_SWEEP, synthetic_batch
    inst = synthetic_batch(MAP50_SWEEP, per, seed=1234, params=p)
S = DsqpSolver(p)
b, _ = S.planes(pack_instances(inst))
os.environ.pop("CSDO_PROFILE", None)
S.refine(b)                       # warm-up (counters are reset when they are read)
os.environ["CSDO_PROFILE"] = "1"
r = S.refine(b)
print(f"admm_iters {int(r.admm_iters.sum())} qps {int(r.n_qp.sum())} n_factor {int(r.n_factor.sum())} launch {S.last_launch()}")
