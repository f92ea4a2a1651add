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


# ---------------------------------------------------------------------------------------------------
# Device-resident agent partition (BASELINE configs[2]): every rank holds the WHOLE batch in HBM (the plane
# build needs every agent's guess), marks its own agents active (csdo_batch.n_active), builds planes for
# and refines only those, and ONE all-gather per step (NCCL on GPUs) gives every rank the whole solution.
def rank_agent_ids(batch: Batch, rank: int, world: int) -> np.ndarray:
    """Global ids of the agents rank `rank` owns: a contiguous slice of every instance's agents (all agents
    of an instance share the horizon, so equal counts are equal work up to the plane counts)."""
    ids: List[int] = []
    for i in range(batch.n_inst):
        a0, a1 = int(batch.inst_agent_ptr[i]), int(batch.inst_agent_ptr[i + 1])
        na = a1 - a0
        ids.extend(range(a0 + (na * rank) // world, a0 + (na * (rank + 1)) // world))
    return np.asarray(ids, np.int32)


class DeviceAllGather:
    """Pack this rank's agents' results -> one all-gather (float64) + one (int32) -> unpack into the full-size
    result tensors, all on the caller's stream; nothing touches the host.  Works on CUDA tensors with NCCL
    and on CPU tensors with gloo (tests)."""

    INT_FIELDS = ("status", "sqp_iters", "n_qp", "admm_iters", "n_factor")

    def __init__(self, batch: Batch, rank: int, world: int, device, dist):
        import torch
        self.rank, self.world, self.dist, self.n_inst = rank, world, dist, batch.n_inst
        off = batch.agent_off
        self.ids, self.ti, self.ci = [], [], []
        for r in range(world):
            ids = rank_agent_ids(batch, r, world)
            t = np.concatenate([np.arange(6 * off[a], 6 * off[a + 1]) for a in ids]) if ids.size else np.zeros(0, np.int64)
            c = np.concatenate([np.arange(8 * off[a], 8 * off[a + 1]) for a in ids]) if ids.size else np.zeros(0, np.int64)
            self.ids.append(torch.from_numpy(ids.astype(np.int64)).to(device))
            self.ti.append(torch.from_numpy(t.astype(np.int64)).to(device))
            self.ci.append(torch.from_numpy(c.astype(np.int64)).to(device))
        self.flen = max(int(self.ti[r].numel() + self.ci[r].numel() + self.ids[r].numel()) for r in range(world))
        self.ilen = max(int(len(self.INT_FIELDS) * self.ids[r].numel()) for r in range(world)) + batch.n_inst
        self.fsend = torch.zeros(self.flen, dtype=torch.float64, device=device)
        self.frecv = torch.zeros(world * self.flen, dtype=torch.float64, device=device)
        self.isend = torch.zeros(self.ilen, dtype=torch.int32, device=device)
        self.irecv = torch.zeros(world * self.ilen, dtype=torch.int32, device=device)
        self.bytes_per_rank = 8 * self.flen + 4 * self.ilen

    def run(self, t: dict) -> None:
        """t: the DeviceResult tensor dict (full-size arrays); call inside the stream context of the refine."""
        import torch
        r = self.rank
        nt, nc, na = self.ti[r].numel(), self.ci[r].numel(), self.ids[r].numel()
        self.fsend[:nt] = t["traj"][self.ti[r]]
        self.fsend[nt:nt + nc] = t["corridors"][self.ci[r]]
        self.fsend[nt + nc:nt + nc + na] = t["objective"][self.ids[r]]
        for j, k in enumerate(self.INT_FIELDS):
            self.isend[j * na:(j + 1) * na] = t[k][self.ids[r]]
        self.isend[self.ilen - self.n_inst:] = t["inst_static_legal"]
        self.dist.all_gather(list(self.frecv.chunk(self.world)), self.fsend)
        self.dist.all_gather(list(self.irecv.chunk(self.world)), self.isend)
        legal = None
        for q in range(self.world):
            f, iv = self.frecv[q * self.flen:(q + 1) * self.flen], self.irecv[q * self.ilen:(q + 1) * self.ilen]
            nt, nc, na = self.ti[q].numel(), self.ci[q].numel(), self.ids[q].numel()
            if q != r:
                t["traj"][self.ti[q]] = f[:nt]
                t["corridors"][self.ci[q]] = f[nt:nt + nc]
                t["objective"][self.ids[q]] = f[nt + nc:nt + nc + na]
                for j, k in enumerate(self.INT_FIELDS):
                    t[k][self.ids[q]] = iv[j * na:(j + 1) * na]
            lg = iv[self.ilen - self.n_inst:]
            legal = lg.clone() if legal is None else torch.minimum(legal, lg)
        t["inst_static_legal"].copy_(legal)   # an instance is legal iff every rank's agents were
