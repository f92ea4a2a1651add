"""Host-side mirror of the reference's DSQP interface on top of the C ABI.

* :class:`DsqpSolver` -- a handle (one GPU, one stream): ``refine``,
  ``refine_device`` (torch tensors already in HBM), ``corridors``, ``planes``.
* :class:`SolverDSQP` -- same constructor arguments, getters and public
  members as the reference class (sqp/dsqp_solver.h:24-47): constructing it
  refines one instance.
* :func:`find_neighbor_pairs_and_planes` -- findNeighborPairsByTrustRegion +
  calcEqualInterPlanes (sqp/inter_agent_cons.h:40-73) for one instance.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import binding
from .batch import Batch, CsdoBatch, CsdoLaunchInfo, CsdoResult, Instance, RefineResult, pack_instances
from .params import CsdoParams, default_params


class DsqpSolver:
    """One csdo_handle: owns a CUDA stream, scratch and the work queue."""

    def __init__(self, params: Optional[CsdoParams] = None, device: int = 0):
        self._lib = binding.lib()
        self.params = params.copy() if params is not None else default_params()
        self.device = device
        h = C.c_void_p()
        rc = self._lib.csdo_create(C.byref(self.params), device, C.byref(h))
        if rc != binding.CSDO_OK:
            raise binding.CsdoError(rc, "csdo_create failed (no usable sm_100 CUDA device?)")
        self._h = h

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.csdo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int) -> None:
        if rc != binding.CSDO_OK:
            raise binding.CsdoError(rc, self._lib.csdo_last_error(self._h).decode())

    def last_launch(self) -> dict:
        info = CsdoLaunchInfo()
        self._lib.csdo_last_launch(self._h, C.byref(info))
        return {n: getattr(info, n) for n, _ in CsdoLaunchInfo._fields_}

    # -- whole refine, host buffers (csdo_refine) --------------------------
    def refine(self, batch: Batch, out: Optional[RefineResult] = None) -> RefineResult:
        res = out if out is not None else RefineResult.allocate(batch)
        cb, cr = batch.to_ctypes(), res.to_ctypes()
        self._check(self._lib.csdo_refine(self._h, C.byref(cb), C.byref(cr)))
        return res

    # -- whole refine, tensors resident in HBM (csdo_refine_device) --------
    def refine_device(self, dbatch: "DeviceBatch", dres: "DeviceResult", stream_ptr: int = 0,
                      by_horizon: bool = True) -> None:
        """Refine a device-resident batch.  by_horizon (default): csdo_refine_device_hinted -- the host copies of
        inst_nt / inst_agent_ptr that the DeviceBatch keeps let the library launch one kernel per horizon class;
        False: csdo_refine_device, one launch shaped for the longest horizon of the batch."""
        sp = C.c_void_p(stream_ptr) if stream_ptr else None
        if not by_horizon:
            self._check(self._lib.csdo_refine_device(self._h, C.byref(dbatch.c), C.byref(dres.c),
                                                     dbatch.max_nt, dbatch.max_planes, sp))
            return
        b = dbatch.host
        nt = np.ascontiguousarray(b.inst_nt, np.int32)
        ptr = np.ascontiguousarray(b.inst_agent_ptr, np.int32)
        ids = getattr(dbatch, "active_ids", None)
        order, n_order = None, 0
        if ids is not None:      # agent-partitioned mode: this rank's agents, longest horizon first
            a_nt = b.agent_nt()[ids]
            o = np.ascontiguousarray(ids[np.argsort(-a_nt, kind="stable")], np.int32)
            order, n_order = o.ctypes.data, int(o.shape[0])
        self._check(self._lib.csdo_refine_device_hinted(self._h, C.byref(dbatch.c), C.byref(dres.c), dbatch.max_planes,
                                                        nt.ctypes.data, ptr.ctypes.data, order, n_order, sp))

    def aggregate_status_device(self, dbatch: "DeviceBatch", dres: "DeviceResult", stream_ptr: int = 0) -> None:
        """SolverDSQP's status aggregation over all agents (after the all-gather of the agent-partitioned mode)."""
        self._check(self._lib.csdo_aggregate_status_device(self._h, C.byref(dbatch.c), C.byref(dres.c),
                                                           C.c_void_p(stream_ptr) if stream_ptr else None))

    def sync(self) -> None:
        """csdo_sync: waits for the last refine_device and raises if the device flagged an error."""
        self._check(self._lib.csdo_sync(self._h))

    # -- pre-process on tensors resident in HBM (csdo_planes_*_device) ------
    def planes_device(self, dbatch: "DeviceBatch", stream_ptr: int = 0, partners: bool = False):
        """findNeighborPairsByTrustRegion + calcEqualInterPlanes on a DeviceBatch: plane_ptr / plane_t /
        plane_abc are created on the device (no host round trip except the plane count) and attached to
        `dbatch`.  Returns (inst_inter_legal tensor, plane_partner tensor or None)."""
        import torch
        b = dbatch.host
        dev = dbatch.t["guess"].device
        steps, A, I = b.total_steps, b.n_agents, b.n_inst
        step_off = torch.empty(steps + 1, dtype=torch.int32, device=dev)
        plane_ptr = torch.zeros(A + 1, dtype=torch.int32, device=dev)
        legal = torch.ones(max(I, 1), dtype=torch.int32, device=dev)
        total = C.c_int64(0)
        sp = C.c_void_p(stream_ptr) if stream_ptr else None
        self._check(self._lib.csdo_planes_count_device(self._h, C.byref(dbatch.c), steps, step_off.data_ptr(),
                                                       plane_ptr.data_ptr(), legal.data_ptr(), C.byref(total), sp))
        n = int(total.value)
        plane_t = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        plane_abc = torch.empty(12 * max(n, 1), dtype=torch.float64, device=dev)
        partner = torch.empty(max(n, 1), dtype=torch.int32, device=dev) if partners else None
        if n:
            self._check(self._lib.csdo_planes_fill_device(self._h, C.byref(dbatch.c), step_off.data_ptr(),
                                                          plane_t.data_ptr(), plane_abc.data_ptr(),
                                                          partner.data_ptr() if partners else None, sp))
        dbatch.set_planes(plane_ptr, plane_t[:n], plane_abc[:12 * n])
        return legal[:I], (partner[:n] if partners else None)

    # -- corridors only (csdo_corridors) -----------------------------------
    def corridors(self, batch: Batch, double_centres: bool = False):
        S = batch.total_steps
        corr = np.zeros(8 * S)
        bs = np.zeros(4 * S, np.int32)
        legal = np.zeros(batch.n_inst, np.int32)
        cb = batch.to_ctypes()
        self._check(self._lib.csdo_corridors(self._h, C.byref(cb), int(double_centres), corr.ctypes.data,
                                             bs.ctypes.data, legal.ctypes.data))
        return corr, bs, legal

    # -- neighbour pairs + planes (csdo_planes_count / csdo_planes_fill) ----
    def planes(self, batch: Batch) -> Tuple[Batch, np.ndarray]:
        ptr = np.zeros(batch.n_agents + 1, np.int32)
        legal = np.zeros(batch.n_inst, np.int32)
        cb = batch.to_ctypes()
        self._check(self._lib.csdo_planes_count(self._h, C.byref(cb), ptr.ctypes.data, legal.ctypes.data))
        total = int(ptr[-1])
        pt = np.zeros(total, np.int32)
        pabc = np.zeros(12 * total)
        if total:
            self._check(self._lib.csdo_planes_fill(self._h, C.byref(cb), ptr.ctypes.data, pt.ctypes.data,
                                                   pabc.ctypes.data))
        return batch.with_planes(ptr, pt, pabc), legal


class DeviceBatch:
    """A Batch whose arrays live in HBM as torch tensors (torch owns the memory)."""

    def __init__(self, batch: Batch, device, order: bool = True):
        import torch
        self.host = batch
        self.t = {}
        for name in ("inst_agent_ptr", "inst_nt", "inst_dims", "obs_ptr", "obs", "agent_off", "guess",
                     "plane_ptr", "plane_t", "plane_abc"):
            self.t[name] = torch.from_numpy(getattr(batch, name)).to(device)
        self.max_nt = int(batch.inst_nt.max()) if batch.n_inst else 0
        k = np.diff(batch.plane_ptr)
        self.max_planes = int(k.max()) if k.size else 0
        if order:
            cost = 13 * batch.agent_nt() + 4 * k.astype(np.int64)
            self.t["agent_order"] = torch.from_numpy(
                np.argsort(-cost, kind="stable").astype(np.int32)).to(device)
        self.c = CsdoBatch()
        self.c.n_inst, self.c.n_agents = batch.n_inst, batch.n_agents
        for name, ten in self.t.items():
            setattr(self.c, name, ten.data_ptr() if ten.numel() else None)
        if not order:
            self.c.agent_order = None

    def h2d_bytes(self) -> int:
        return int(sum(t.numel() * t.element_size() for t in self.t.values()))

    def set_active(self, agent_ids: np.ndarray) -> None:
        """Agent-partitioned mode: only these agents (global ids) get planes and are refined
        (csdo_batch.agent_order + n_active)."""
        import torch
        self.active_ids = np.ascontiguousarray(agent_ids, np.int32)
        self.t["agent_order"] = torch.from_numpy(self.active_ids).to(self.t["guess"].device)
        self.c.agent_order = self.t["agent_order"].data_ptr()
        self.c.n_active = int(self.active_ids.shape[0])

    def set_planes(self, plane_ptr, plane_t, plane_abc) -> None:
        """Attach device-resident planes (DsqpSolver.planes_device) and the processing order they imply."""
        import torch
        self.t["plane_ptr"], self.t["plane_t"], self.t["plane_abc"] = plane_ptr, plane_t, plane_abc
        k = torch.diff(plane_ptr.to(torch.int64))
        self.max_planes = int(k.max().item()) if k.numel() else 0
        nt = torch.from_numpy(self.host.agent_nt()).to(plane_ptr.device)
        cost = 13 * nt + 4 * k
        active = getattr(self, "active_ids", None)
        if active is None:
            self.t["agent_order"] = torch.argsort(-cost, stable=True).to(torch.int32)
        else:   # longest first among the rank's own agents
            ids = torch.from_numpy(active.astype(np.int64)).to(plane_ptr.device)
            self.t["agent_order"] = ids[torch.argsort(-cost[ids], stable=True)].to(torch.int32)
            self.c.n_active = int(active.shape[0])
        for name in ("plane_ptr", "plane_t", "plane_abc", "agent_order"):
            ten = self.t[name]
            setattr(self.c, name, ten.data_ptr() if ten.numel() else None)

    def planes_to_host(self) -> Batch:
        """The batch with the device-built planes copied back (for checks against the oracle)."""
        return self.host.with_planes(self.t["plane_ptr"].cpu().numpy(), self.t["plane_t"].cpu().numpy(),
                                     self.t["plane_abc"].cpu().numpy())


class DeviceResult:
    def __init__(self, batch: Batch, device):
        import torch
        A, S, I = batch.n_agents, batch.total_steps, batch.n_inst
        f = lambda n: torch.zeros(n, dtype=torch.float64, device=device)
        i = lambda n: torch.zeros(n, dtype=torch.int32, device=device)
        self.t = dict(traj=f(6 * S), corridors=f(8 * S), status=i(A), sqp_iters=i(A), n_qp=i(A),
                      admm_iters=i(A), n_factor=i(A), objective=f(A), inst_status=i(I),
                      inst_static_legal=i(I))
        self.c = CsdoResult()
        for name, ten in self.t.items():
            setattr(self.c, name, ten.data_ptr() if ten.numel() else None)

    def to_host(self) -> RefineResult:
        return RefineResult(**{k: v.cpu().numpy() for k, v in self.t.items()})

    def counters_to_host(self) -> dict:
        """Only the small per-agent / per-instance arrays (no trajectories)."""
        return {k: v.cpu().numpy() for k, v in self.t.items() if k not in ("traj", "corridors")}


# ---------------------------------------------------------------------------
# reference-shaped interface
@dataclass
class OptimizeResult:   # sqp/common.h:14-22
    x: float = 0.0
    y: float = 0.0
    yaw: float = 0.0
    v: float = 0.0
    a: float = 0.0
    steer: float = 0.0
    d_steer: float = 0.0


@dataclass
class Corridor:         # sqp/corridor.h:8-11
    xf_min: float
    xf_max: float
    yf_min: float
    yf_max: float
    xr_min: float
    xr_max: float
    yr_min: float
    yr_max: float


def _guess_planes(x0_bar: Sequence[Sequence[OptimizeResult]]) -> np.ndarray:
    na, nt = len(x0_bar), len(x0_bar[0])
    g = np.zeros((na, 6, nt))
    for a, row in enumerate(x0_bar):
        assert len(row) == nt, "all agents must share the horizon"
        for t, r in enumerate(row):
            g[a, :, t] = (r.x, r.y, r.yaw, r.steer, r.v, r.d_steer)
    return g


def find_neighbor_pairs_and_planes(x0_bar, params: Optional[CsdoParams] = None,
                                   solver: Optional[DsqpSolver] = None):
    """-> (inter_planes: per agent list of (t, 12 coefficients), initial_inter_legal)."""
    own = solver is None
    solver = solver or DsqpSolver(params)
    try:
        g = _guess_planes(x0_bar) if not isinstance(x0_bar, np.ndarray) else x0_bar
        b = pack_instances([Instance(g, 1.0, 1.0, np.zeros((0, 3)))])
        pb, legal = solver.planes(b)
        planes = []
        for a in range(pb.n_agents):
            k0, k1 = int(pb.plane_ptr[a]), int(pb.plane_ptr[a + 1])
            planes.append((pb.plane_t[k0:k1].copy(), pb.plane_abc[12 * k0:12 * k1].reshape(-1, 12).copy()))
        return planes, bool(legal[0])
    finally:
        if own:
            solver.close()


class SolverDSQP:
    """Mirror of the reference class: the constructor refines one instance.

    SolverDSQP(solutions, x0_bar, inter_planes, dimx, dimy, obstacles, param, logger_level)
    (sqp/dsqp_solver.h:26-34).  ``solutions`` is filled in place (a list that
    receives one list of OptimizeResult per agent).  ``inter_planes`` is a list
    per agent of (plane_t, plane_abc[K,12]).  ``obstacles`` is an iterable of
    (x, y, r) in the caller's container order.
    """

    def __init__(self, solutions: list, x0_bar, inter_planes, dimx: float, dimy: float, obstacles,
                 param: Optional[CsdoParams] = None, logger_level: int = 2,
                 solver: Optional[DsqpSolver] = None):
        g = _guess_planes(x0_bar) if not isinstance(x0_bar, np.ndarray) else np.asarray(x0_bar, np.float64)
        obs = np.asarray([tuple(o) for o in obstacles], np.float64).reshape(-1, 3)
        ins = Instance(g, dimx, dimy, obs, [np.asarray(p[0], np.int32) for p in inter_planes],
                       [np.asarray(p[1], np.float64).reshape(-1, 12) for p in inter_planes])
        batch = pack_instances([ins])
        own = solver is None
        solver = solver or DsqpSolver(param)
        try:
            import time
            t0 = time.perf_counter()
            res = solver.refine(batch)
            self._runtime = time.perf_counter() - t0
        finally:
            if own:
                solver.close()
        na, nt = g.shape[0], g.shape[2]
        self.result = res
        self.num_iterations: List[int] = [int(v) for v in res.sqp_iters]
        self.corridors: List[List[Corridor]] = []
        solutions.clear()
        for a in range(na):
            tr, co = res.agent_traj(batch, a), res.agent_corridor(batch, a)
            solutions.append([OptimizeResult(x=tr[0, t], y=tr[1, t], yaw=tr[2, t], steer=tr[3, t],
                                             v=tr[4, t], d_steer=tr[5, t]) for t in range(nt)])
            self.corridors.append([Corridor(*co[:, t]) for t in range(nt)])
        self._status = int(res.inst_status[0])
        self._static_legal = bool(res.inst_static_legal[0])

    def getSolverStatus(self) -> int:
        return self._status

    def getMaxOfRuntimes(self) -> float:
        """Reference: max per-agent time + bookkeeping; here all agents run concurrently."""
        return self._runtime

    def get_initial_static_legal(self) -> bool:
        return self._static_legal
