"""Developer tooling: device-resident refine timing of one workload shape (not a bench line).

usage: dev_shape.py <c5|c5a|c5b|c5c|map50|room> <instances per shape> [reps]
"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from csdotrajectoryplanning_b200 import default_params, pack_instances
from csdotrajectoryplanning_b200.solver import DeviceBatch, DeviceResult, DsqpSolver
from tools import synth
import bench

name = sys.argv[1]
per = int(sys.argv[2]) if len(sys.argv) > 2 else 4
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
p = default_params()
if name.startswith("c5"):
    shapes = {"c5": synth.C5_SHAPES, "c5a": synth.C5_SHAPES[:1], "c5b": synth.C5_SHAPES[1:2], "c5c": synth.C5_SHAPES[2:]}[name]
    inst = synth.synth_batch(shapes, per, 1234, p)
elif name == "map50":
    from csdotrajectoryplanning_b200.scenario import MAP50_SWEEP, synthetic_batch
    inst = synthetic_batch(MAP50_SWEEP, per, seed=1234, params=p)
elif name == "real":
    inst = bench.build_instances("real", per, 0, 1, p)
else:
    raise SystemExit("unknown workload")
S = DsqpSolver(p)
b, _ = S.planes(pack_instances(inst))
dev = torch.device("cuda", 0)
db, dr = DeviceBatch(b, dev), DeviceResult(b, dev)
stream = torch.cuda.Stream(device=dev)
ts = []
for r in range(reps + 1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream); S.refine_device(db, dr, stream.cuda_stream); e1.record(stream)
    torch.cuda.synchronize()
    if r: ts.append(e0.elapsed_time(e1) * 1e-3)
res = dr.to_host()
dt = min(ts)
fl = bench.algorithmic_flops(b, res)
print(json.dumps({"workload": name, "instances": len(inst), "agents": int(b.n_agents), "horizon_max": int(b.inst_nt.max()),
                  "planes": int(b.plane_ptr[-1]), "max_planes_per_agent": int(np.diff(b.plane_ptr).max()),
                  "qps": int(res.n_qp.sum()), "admm_iters": int(res.admm_iters.sum()), "n_factor": int(res.n_factor.sum()),
                  "seconds": dt, "qp_per_s": float(res.n_qp.sum() / dt), "admm_iter_per_s": float(res.admm_iters.sum() / dt),
                  "tflops": fl / dt * 1e-12, "refine_ms_per_instance": 1e3 * dt / len(inst),
                  "status_hist": {int(k): int(v) for k, v in zip(*np.unique(res.status, return_counts=True))},
                  "launch": S.last_launch()}))
if os.environ.get("CSDO_PROFILE"):   # -DCSDO_DEV_TIMERS build: the host-pointer call prints the phase timers
    S.refine(b)
