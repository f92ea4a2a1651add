"""Agent-partitioned refine of 100-agent instances over N GPUs with ONE NCCL all-gather (BASELINE config 3 shape).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/agent_partition_demo.py

Every rank refines its slice of each instance's agents on its own GPU (planes were built from the full
guess, so results equal the unsharded run), then all ranks all-gather trajectories/statuses over NCCL.
Rank 0 also runs the unsharded refine and checks bit-equality.
"""
import json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from csdotrajectoryplanning_b200 import default_params, pack_instances, sharding
from csdotrajectoryplanning_b200.scenario import synthetic_batch, MAP100_A100
from csdotrajectoryplanning_b200.solver import DsqpSolver

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
p = default_params()
n_inst = int(os.environ.get("CSDO_DEMO_INSTANCES", "2"))
inst = synthetic_batch(MAP100_A100, n_inst, seed=4242, params=p)     # same seed on every rank
S = DsqpSolver(p, device=lr)
batch, _ = S.planes(pack_instances(inst))                             # planes from the FULL guess
torch.cuda.synchronize()
if world > 1: dist.barrier()
t0 = time.perf_counter()
full = sharding.refine_agent_partitioned(batch, S.refine, rank, world, dist if world > 1 else None, dev)
torch.cuda.synchronize()
if world > 1: dist.barrier()
dt = time.perf_counter() - t0
if rank == 0:
    whole = S.refine(batch)
    same = all(np.array_equal(getattr(full, k), getattr(whole, k)) for k in
               ("traj", "corridors", "status", "sqp_iters", "admm_iters", "n_factor", "inst_status", "inst_static_legal"))
    print(json.dumps({"world": world, "instances": batch.n_inst, "agents": batch.n_agents, "horizon": int(batch.inst_nt.max()),
                      "planes": int(batch.plane_ptr[-1]), "qps": int(full.n_qp.sum()), "seconds_partitioned_incl_allgather": dt,
                      "bit_identical_to_unsharded": bool(same), "status_hist": np.unique(full.status, return_counts=True)[0].tolist()}))
if world > 1:
    dist.destroy_process_group()
