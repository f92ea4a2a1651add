"""Developer tooling: BASELINE configs[2]-shaped workload (map100by100, 100 agents, 50 obstacles) timing."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
from csdotrajectoryplanning_b200 import default_params, pack_instances
from csdotrajectoryplanning_b200.scenario import MAP100_A100, synthetic_batch
from csdotrajectoryplanning_b200.solver import DsqpSolver

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
p = default_params()
t = time.time(); inst = synthetic_batch(MAP100_A100, n, seed=4321, params=p); tg = time.time() - t
S = DsqpSolver(p)
b, _ = S.planes(pack_instances(inst))
S.refine(b)
ts = []
for _ in range(3):
    t = time.time(); r = S.refine(b); ts.append(time.time() - t)
dt = min(ts)
print(json.dumps({"workload": f"map100by100-shaped: {n} instances x 100 agents, 50 obstacles", "agents": int(b.n_agents),
                  "horizon_max": int(b.inst_nt.max()), "planes": int(b.plane_ptr[-1]), "qps": int(r.n_qp.sum()),
                  "admm_iters": int(r.admm_iters.sum()), "seconds_e2e": dt, "qp_per_s": float(r.n_qp.sum() / dt),
                  "refine_ms_per_instance": 1e3 * dt / n, "status_hist": np.bincount(np.abs(r.inst_status)).tolist(),
                  "launch": S.last_launch(), "generation_s": tg}))
