"""Batch containers: the flat layout of include/csdo_dsqp.h as numpy arrays.

An :class:`Instance` is what ``csdo.cc:113-147`` hands to ``SolverDSQP``:
the interpolated guess ``x0_bar`` (Na x Nt), the inter-vehicle planes, the
map size and the obstacle list.  A :class:`Batch` is many instances
flattened (agents numbered globally, CSR pointers per instance/agent).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np


class CsdoBatch(C.Structure):
    _fields_ = [
        ("n_inst", C.c_int32), ("n_agents", C.c_int32),
        ("inst_agent_ptr", C.c_void_p), ("inst_nt", C.c_void_p),
        ("inst_dims", C.c_void_p), ("obs_ptr", C.c_void_p), ("obs", C.c_void_p),
        ("agent_off", C.c_void_p), ("guess", C.c_void_p),
        ("plane_ptr", C.c_void_p), ("plane_t", C.c_void_p), ("plane_abc", C.c_void_p),
        ("agent_order", C.c_void_p), ("n_active", C.c_int32),
    ]


class CsdoResult(C.Structure):
    _fields_ = [
        ("traj", C.c_void_p), ("corridors", C.c_void_p), ("status", C.c_void_p),
        ("sqp_iters", C.c_void_p), ("n_qp", C.c_void_p), ("admm_iters", C.c_void_p),
        ("n_factor", C.c_void_p), ("objective", C.c_void_p),
        ("inst_status", C.c_void_p), ("inst_static_legal", C.c_void_p),
    ]


class CsdoLaunchInfo(C.Structure):
    _fields_ = [("launches", C.c_int32), ("grid", C.c_int32), ("block", C.c_int32),
                ("smem_bytes", C.c_int32), ("tier", C.c_int32), ("ctas_per_sm", C.c_int32)]


@dataclass
class Instance:
    """One planning instance at the SolverDSQP boundary.

    guess: (Na, 6, Nt) float64 planes x,y,yaw,steer,v,w (OptimizeResult fields,
    sqp/common.h:14-22; w == d_steer).  obstacles: (No, 3) x,y,r in container
    order.  plane_t/plane_abc: per-agent lists (None until planes are built).
    """
    guess: np.ndarray
    dimx: float
    dimy: float
    obstacles: np.ndarray
    plane_t: Optional[List[np.ndarray]] = None
    plane_abc: Optional[List[np.ndarray]] = None
    name: str = ""

    @property
    def n_agents(self) -> int:
        return int(self.guess.shape[0])

    @property
    def nt(self) -> int:
        return int(self.guess.shape[2])


@dataclass
class Batch:
    inst_agent_ptr: np.ndarray
    inst_nt: np.ndarray
    inst_dims: np.ndarray
    obs_ptr: np.ndarray
    obs: np.ndarray
    agent_off: np.ndarray
    guess: np.ndarray
    plane_ptr: np.ndarray
    plane_t: np.ndarray
    plane_abc: np.ndarray
    names: List[str] = field(default_factory=list)
    agent_order: Optional[np.ndarray] = None   # int32 processing order; with n_active > 0: the active subset
    n_active: int = 0

    @property
    def n_inst(self) -> int:
        return int(self.inst_nt.shape[0])

    @property
    def n_agents(self) -> int:
        return int(self.agent_off.shape[0] - 1)

    @property
    def total_steps(self) -> int:
        return int(self.agent_off[-1])

    def agent_nt(self) -> np.ndarray:
        return np.diff(self.agent_off).astype(np.int64)

    def agent_guess(self, a: int) -> np.ndarray:
        o, nt = int(self.agent_off[a]), int(self.agent_off[a + 1] - self.agent_off[a])
        return self.guess[6 * o: 6 * (o + nt)].reshape(6, nt)

    def validate(self) -> None:
        assert self.inst_agent_ptr.dtype == np.int32 and self.inst_nt.dtype == np.int32
        assert self.agent_off.dtype == np.int64 and self.plane_ptr.dtype == np.int32
        assert self.inst_agent_ptr[0] == 0 and self.inst_agent_ptr[-1] == self.n_agents
        nt = np.repeat(self.inst_nt, np.diff(self.inst_agent_ptr))
        assert np.array_equal(nt.astype(np.int64), self.agent_nt()), "agent_off != inst_nt"
        assert self.guess.shape == (6 * self.total_steps,)
        assert self.plane_t.shape[0] == self.plane_ptr[-1]
        assert self.plane_abc.shape == (int(self.plane_ptr[-1]) * 12,)

    def with_planes(self, plane_ptr, plane_t, plane_abc) -> "Batch":
        return Batch(self.inst_agent_ptr, self.inst_nt, self.inst_dims, self.obs_ptr, self.obs,
                     self.agent_off, self.guess, np.ascontiguousarray(plane_ptr, np.int32),
                     np.ascontiguousarray(plane_t, np.int32),
                     np.ascontiguousarray(plane_abc, np.float64).reshape(-1), list(self.names))

    def to_ctypes(self) -> CsdoBatch:
        """ctypes view; the Batch must outlive the returned struct."""
        b = CsdoBatch()
        b.n_inst, b.n_agents = self.n_inst, self.n_agents
        for name in ("inst_agent_ptr", "inst_nt", "inst_dims", "obs_ptr", "obs", "agent_off",
                     "guess", "plane_ptr", "plane_t", "plane_abc"):
            arr = getattr(self, name)
            assert arr.flags["C_CONTIGUOUS"]
            setattr(b, name, arr.ctypes.data if arr.size else None)
        if self.agent_order is not None:
            assert self.agent_order.dtype == np.int32 and self.agent_order.flags["C_CONTIGUOUS"]
            b.agent_order = self.agent_order.ctypes.data
            b.n_active = int(self.n_active)
        else:
            b.agent_order = None
            b.n_active = 0
        return b

    def select_instances(self, idx: Sequence[int]) -> "Batch":
        """Sub-batch of whole instances (used for instance sharding)."""
        return pack_instances(self.unpack(idx))

    def unpack(self, idx: Optional[Sequence[int]] = None) -> List[Instance]:
        out = []
        for i in (range(self.n_inst) if idx is None else idx):
            a0, a1 = int(self.inst_agent_ptr[i]), int(self.inst_agent_ptr[i + 1])
            nt = int(self.inst_nt[i])
            g = np.stack([self.agent_guess(a) for a in range(a0, a1)]) if a1 > a0 \
                else np.zeros((0, 6, nt))
            pts, pabc = [], []
            for a in range(a0, a1):
                k0, k1 = int(self.plane_ptr[a]), int(self.plane_ptr[a + 1])
                pts.append(self.plane_t[k0:k1].copy())
                pabc.append(self.plane_abc[12 * k0:12 * k1].reshape(-1, 12).copy())
            o0, o1 = int(self.obs_ptr[i]), int(self.obs_ptr[i + 1])
            out.append(Instance(g, float(self.inst_dims[2 * i]), float(self.inst_dims[2 * i + 1]),
                                self.obs[3 * o0:3 * o1].reshape(-1, 3).copy(), pts, pabc,
                                self.names[i] if i < len(self.names) else ""))
        return out


def pack_instances(instances: Sequence[Instance]) -> Batch:
    n_inst = len(instances)
    inst_agent_ptr = np.zeros(n_inst + 1, np.int32)
    inst_nt = np.zeros(n_inst, np.int32)
    inst_dims = np.zeros(2 * n_inst, np.float64)
    obs_ptr = np.zeros(n_inst + 1, np.int32)
    obs_l, guess_l, off_l, pt_l, pabc_l, pcnt = [], [], [0], [], [], [0]
    for i, ins in enumerate(instances):
        g = np.ascontiguousarray(ins.guess, np.float64)
        assert g.ndim == 3 and g.shape[1] == 6, "guess must be (Na, 6, Nt)"
        na, nt = g.shape[0], g.shape[2]
        inst_agent_ptr[i + 1] = inst_agent_ptr[i] + na
        inst_nt[i] = nt
        inst_dims[2 * i], inst_dims[2 * i + 1] = ins.dimx, ins.dimy
        ob = np.ascontiguousarray(ins.obstacles, np.float64).reshape(-1, 3)
        obs_ptr[i + 1] = obs_ptr[i] + ob.shape[0]
        obs_l.append(ob.reshape(-1))
        for a in range(na):
            guess_l.append(g[a].reshape(-1))
            off_l.append(off_l[-1] + nt)
            if ins.plane_t is not None:
                t = np.ascontiguousarray(ins.plane_t[a], np.int32)
                pt_l.append(t)
                pabc_l.append(np.ascontiguousarray(ins.plane_abc[a], np.float64).reshape(-1))
                pcnt.append(pcnt[-1] + t.shape[0])
            else:
                pcnt.append(pcnt[-1])
    cat = lambda l, dt: (np.concatenate(l).astype(dt) if l else np.zeros(0, dt))
    return Batch(inst_agent_ptr, inst_nt, inst_dims, obs_ptr, cat(obs_l, np.float64),
                 np.asarray(off_l, np.int64), cat(guess_l, np.float64),
                 np.asarray(pcnt, np.int32), cat(pt_l, np.int32), cat(pabc_l, np.float64),
                 [ins.name for ins in instances])


@dataclass
class RefineResult:
    traj: np.ndarray
    corridors: np.ndarray
    status: np.ndarray
    sqp_iters: np.ndarray
    n_qp: np.ndarray
    admm_iters: np.ndarray
    n_factor: np.ndarray
    objective: np.ndarray
    inst_status: np.ndarray
    inst_static_legal: np.ndarray

    @staticmethod
    def allocate(batch: Batch) -> "RefineResult":
        A, S, I = batch.n_agents, batch.total_steps, batch.n_inst
        z = lambda n: np.zeros(n, np.int32)
        return RefineResult(np.zeros(6 * S), np.zeros(8 * S), z(A), z(A), z(A), z(A), z(A),
                            np.zeros(A), z(I), z(I))

    def to_ctypes(self) -> CsdoResult:
        r = CsdoResult()
        for name, _ in CsdoResult._fields_:
            arr = getattr(self, name)
            setattr(r, name, arr.ctypes.data if arr.size else None)
        return r

    def agent_traj(self, batch: Batch, a: int) -> np.ndarray:
        o, nt = int(batch.agent_off[a]), int(batch.agent_off[a + 1] - batch.agent_off[a])
        return self.traj[6 * o:6 * (o + nt)].reshape(6, nt)

    def agent_corridor(self, batch: Batch, a: int) -> np.ndarray:
        o, nt = int(batch.agent_off[a]), int(batch.agent_off[a + 1] - batch.agent_off[a])
        return self.corridors[8 * o:8 * (o + nt)].reshape(8, nt)
