"""Multi-GPU sharding of the DSQP refine stage (one process per GPU).

Instances are independent, and inside an instance agents are independent for
the whole refine because the inter-vehicle planes are frozen at the initial
guess (reference csdo.cc:119-129; SolverDSQP only reads them,
sqp/dsqp_solver.cc:153).  So:

* instance sharding (:func:`shard_instances`): contiguous instance ranges per
  rank balanced by estimated work; NO data-path collective;
* agent partitioning of one large instance (:func:`shard_agents`): each rank
  refines the agents [a0, a1) of every instance; the planes of those agents
  were built from the full guess, so results are identical to the unsharded
  run.  The only exchange is one all-gather of the final trajectories and
  statuses (:func:`allgather_agent_results`) so that every rank holds the
  whole solution (the reference's single-process output), over NCCL on GPUs
  (gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import numpy as np

from .batch import Batch, RefineResult, pack_instances


def agent_cost(batch: Batch) -> np.ndarray:
    """Rows of the agent QP, the unit the kernel's work scales with (m = 13 Nt + 4 K)."""
    return 13 * batch.agent_nt() + 4 * np.diff(batch.plane_ptr).astype(np.int64)


def split_balanced(cost: np.ndarray, parts: int) -> List[Tuple[int, int]]:
    """Contiguous ranges with near-equal total cost."""
    n = int(cost.shape[0])
    if parts <= 1 or n == 0:
        return [(0, n)] + [(n, n)] * (max(parts, 1) - 1)
    csum = np.concatenate([[0], np.cumsum(cost, dtype=np.float64)])
    bounds = [0]
    for r in range(1, parts):
        target = csum[-1] * r / parts
        j = int(np.searchsorted(csum, target, side="left"))
        if j > 0 and j <= n and target - csum[j - 1] <= csum[j] - target:
            j -= 1
        bounds.append(min(max(j, bounds[-1]), n))
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(parts)]


def shard_instances(batch: Batch, rank: int, world: int) -> Tuple[Batch, Tuple[int, int]]:
    """Sub-batch of whole instances for `rank`; returns (batch, (i0, i1))."""
    ac = agent_cost(batch)
    inst_cost = np.add.reduceat(ac, batch.inst_agent_ptr[:-1].astype(np.int64)) if batch.n_agents else ac
    if batch.n_inst and batch.n_agents:
        empty = np.diff(batch.inst_agent_ptr) == 0
        inst_cost = np.where(empty, 0, inst_cost)
    i0, i1 = split_balanced(np.asarray(inst_cost), world)[rank]
    return batch.select_instances(range(i0, i1)), (i0, i1)


def shard_agents(batch: Batch, rank: int, world: int) -> Tuple[Batch, np.ndarray]:
    """Agent partition: every instance keeps its obstacles/dims, each rank gets a
    contiguous slice of each instance's agents.  Returns (sub-batch, global agent ids)."""
    insts = batch.unpack()
    ids: List[int] = []
    sub = []
    for i, ins in enumerate(insts):
        a0 = int(batch.inst_agent_ptr[i])
        ac = agent_cost(pack_instances([ins])) if ins.n_agents else np.zeros(0, np.int64)
        lo, hi = split_balanced(ac, world)[rank]
        ids.extend(range(a0 + lo, a0 + hi))
        from .batch import Instance
        sub.append(Instance(ins.guess[lo:hi], ins.dimx, ins.dimy, ins.obstacles,
                            ins.plane_t[lo:hi], ins.plane_abc[lo:hi], ins.name))
    return pack_instances(sub), np.asarray(ids, np.int64)


def scatter_agent_results(full: RefineResult, batch: Batch, ids: np.ndarray, part: RefineResult,
                          part_batch: Batch) -> None:
    """Write a rank's agent results into the full-size result arrays."""
    for j, a in enumerate(ids):
        o, nt = int(batch.agent_off[a]), int(batch.agent_off[a + 1] - batch.agent_off[a])
        po = int(part_batch.agent_off[j])
        full.traj[6 * o:6 * (o + nt)] = part.traj[6 * po:6 * (po + nt)]
        full.corridors[8 * o:8 * (o + nt)] = part.corridors[8 * po:8 * (po + nt)]
        for k in ("status", "sqp_iters", "n_qp", "admm_iters", "n_factor", "objective"):
            getattr(full, k)[a] = getattr(part, k)[j]


def aggregate_instance_status(batch: Batch, res: RefineResult) -> None:
    """SolverDSQP's aggregation (dsqp_solver.cc:1224-1243) over ALL agents of each instance."""
    for i in range(batch.n_inst):
        worst, anyu = 2, False
        for a in range(int(batch.inst_agent_ptr[i]), int(batch.inst_agent_ptr[i + 1])):
            s = int(res.status[a])
            if abs(s) > 1:
                anyu = True
                if abs(s) > worst:
                    worst = s
        res.inst_status[i] = worst if anyu else 1


def refine_agent_partitioned(batch: Batch, refine_fn: Callable[[Batch], RefineResult], rank: int,
                             world: int, dist=None, device=None) -> RefineResult:
    """Agent-partitioned refine with ONE all-gather of the results.

    refine_fn: Batch -> RefineResult (the CUDA path in production).  dist: the
    torch.distributed module (initialised) or None for world == 1.
    """
    import torch
    part_batch, ids = shard_agents(batch, rank, world)
    part = refine_fn(part_batch)
    full = RefineResult.allocate(batch)
    if world == 1 or dist is None:
        scatter_agent_results(full, batch, ids, part, part_batch)
        full.inst_static_legal[:] = part.inst_static_legal
        aggregate_instance_status(batch, full)
        return full
    # fixed-size record per rank: pad to the largest shard so a plain all_gather works
    shards = [shard_agents(batch, r, world) for r in range(world)]
    max_steps = max(int(b.total_steps) for b, _ in shards)
    max_agents = max(int(b.n_agents) for b, _ in shards)
    rec = torch.zeros(14 * max_steps + 6 * max_agents + batch.n_inst, dtype=torch.float64)
    s, a = part_batch.total_steps, part_batch.n_agents
    rec[:6 * s] = torch.from_numpy(part.traj)
    rec[6 * max_steps:6 * max_steps + 8 * s] = torch.from_numpy(part.corridors)
    base = 14 * max_steps
    for j, k in enumerate(("status", "sqp_iters", "n_qp", "admm_iters", "n_factor", "objective")):
        rec[base + j * max_agents: base + j * max_agents + a] = torch.from_numpy(
            np.asarray(getattr(part, k), np.float64))
    rec[base + 6 * max_agents:] = torch.from_numpy(part.inst_static_legal.astype(np.float64))
    if device is not None:
        rec = rec.to(device)
    out = [torch.empty_like(rec) for _ in range(world)]
    dist.all_gather(out, rec)
    legal = np.ones(batch.n_inst, np.int32)
    for r in range(world):
        o = out[r].cpu().numpy()
        pb, pids = shards[r]
        s, a = pb.total_steps, pb.n_agents
        pr = RefineResult.allocate(pb)
        pr.traj[:] = o[:6 * s]
        pr.corridors[:] = o[6 * max_steps:6 * max_steps + 8 * s]
        for j, k in enumerate(("status", "sqp_iters", "n_qp", "admm_iters", "n_factor", "objective")):
            v = o[base + j * max_agents: base + j * max_agents + a]
            getattr(pr, k)[:] = v if k == "objective" else np.rint(v).astype(np.int32)
        scatter_agent_results(full, batch, pids, pr, pb)
        legal &= np.rint(o[base + 6 * max_agents:]).astype(np.int32)
    full.inst_static_legal[:] = legal
    aggregate_instance_status(batch, full)
    return full
