"""TEST INFRASTRUCTURE -- ctypes loader of the CPU oracle (oracle/_build).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs import this module.  PARITY UNPINNED: the reference
tree holds no golden vector for the DSQP path and neither the reference nor
OSQP 0.6.3 can be built offline (see oracle/dsqp_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from csdotrajectoryplanning_b200.batch import Batch, CsdoBatch, CsdoResult, RefineResult
from csdotrajectoryplanning_b200.params import CsdoParams

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libdsqp_oracle.so")
_lib = None

_dp = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_generate_box.argtypes = [C.POINTER(CsdoParams), C.c_double, C.c_double, C.c_double,
                                       C.c_double, C.c_void_p, C.c_int, _dp, _ip]
        L.orc_generate_box.restype = None
        L.orc_agent_corridors.argtypes = [C.POINTER(CsdoParams), C.c_int, _dp, _dp, _dp, C.c_double,
                                          C.c_double, C.c_void_p, C.c_int, C.c_int, _dp, C.c_void_p]
        L.orc_agent_corridors.restype = C.c_int
        L.orc_instance_planes.argtypes = [C.POINTER(CsdoParams), C.c_int, C.c_int, _dp, _ip,
                                          C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_instance_planes.restype = C.c_int
        L.orc_assemble_qp.argtypes = [C.POINTER(CsdoParams), C.c_int, _dp, _dp, _dp, _dp, C.c_int,
                                      C.c_void_p, C.c_void_p, _ip, _ip, _dp, _dp, _dp, _ip, _ip, _dp]
        L.orc_assemble_qp.restype = C.c_int
        L.orc_osqp_solve.argtypes = [C.c_int, C.c_int, _ip, _ip, _dp, _dp, _ip, _ip, _dp, _dp, _dp,
                                     C.c_void_p, C.POINTER(CsdoParams), C.c_int, C.c_int, _dp,
                                     C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                     C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.orc_osqp_solve.restype = C.c_int
        L.orc_refine.argtypes = [C.POINTER(CsdoParams), C.POINTER(CsdoBatch), C.POINTER(CsdoResult),
                                 C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.orc_refine.restype = C.c_int
        L.orc_interpolate_guess.argtypes = [C.c_int, _dp, _ip, C.c_void_p, C.c_int, C.c_double,
                                            C.c_double, C.c_double, C.c_double, C.c_int, _dp]
        L.orc_interpolate_guess.restype = C.c_int
        _lib = L
    return _lib


def _vp(a):
    return a.ctypes.data if a is not None and a.size else None


def generate_box(p: CsdoParams, dimx, dimy, x, y, obs: np.ndarray):
    obs = np.ascontiguousarray(obs, np.float64).reshape(-1, 3)
    box = np.zeros(4)
    st = np.zeros(2, np.int32)
    lib().orc_generate_box(C.byref(p), dimx, dimy, x, y, _vp(obs), obs.shape[0], box, st)
    return box, st


def agent_corridors(p: CsdoParams, x, y, yaw, dimx, dimy, obs, double_centres: bool):
    nt = x.shape[0]
    obs = np.ascontiguousarray(obs, np.float64).reshape(-1, 3)
    corr = np.zeros(8 * nt)
    bs = np.zeros(4 * nt, np.int32)
    legal = lib().orc_agent_corridors(C.byref(p), nt, np.ascontiguousarray(x), np.ascontiguousarray(y),
                                      np.ascontiguousarray(yaw), dimx, dimy, _vp(obs), obs.shape[0],
                                      int(double_centres), corr, bs.ctypes.data)
    return corr.reshape(8, nt), bs.reshape(nt, 2, 2), bool(legal)


def instance_planes(p: CsdoParams, guess: np.ndarray):
    """guess (Na,6,Nt) -> (per-agent plane_t list, plane_abc list, inter_legal)."""
    na, _, nt = guess.shape
    g = np.ascontiguousarray(guess, np.float64).reshape(-1)
    cnt = np.zeros(na, np.int32)
    lib().orc_instance_planes(C.byref(p), na, nt, g, cnt, None, None, None)
    ptr = np.zeros(na + 1, np.int32)
    ptr[1:] = np.cumsum(cnt)
    pt = np.zeros(int(ptr[-1]), np.int32)
    pabc = np.zeros(int(ptr[-1]) * 12)
    cnt2 = np.zeros(na, np.int32)
    legal = lib().orc_instance_planes(C.byref(p), na, nt, g, cnt2, _vp(pt) or pt.ctypes.data,
                                      _vp(pabc) or pabc.ctypes.data, ptr.ctypes.data)
    pts = [pt[ptr[a]:ptr[a + 1]].copy() for a in range(na)]
    pabcs = [pabc[12 * ptr[a]:12 * ptr[a + 1]].reshape(-1, 12).copy() for a in range(na)]
    return pts, pabcs, bool(legal)


def assemble_qp(p: CsdoParams, lin, trust, cfg, corr, plane_t, plane_abc):
    """-> scipy-free CSC arrays of the agent QP in the reference's ordering."""
    nt = lin.shape[1]
    K = 0 if plane_t is None else int(plane_t.shape[0])
    n, m, nnz = 6 * nt - 2, 13 * nt + 4 * K, 28 * nt - 11 + 12 * K
    Ap, Ai, Ax = np.zeros(n + 1, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz)
    l, u = np.zeros(m), np.zeros(m)
    Pp, Pi, Px = np.zeros(n + 1, np.int32), np.zeros(3 * nt, np.int32), np.zeros(3 * nt)
    pt = np.ascontiguousarray(plane_t, np.int32) if K else None
    pa = np.ascontiguousarray(plane_abc, np.float64).reshape(-1) if K else None
    npz = lib().orc_assemble_qp(C.byref(p), nt, np.ascontiguousarray(lin).reshape(-1),
                                np.ascontiguousarray(trust).reshape(-1), np.ascontiguousarray(cfg),
                                np.ascontiguousarray(corr).reshape(-1), K, _vp(pt), _vp(pa),
                                Ap, Ai, Ax, l, u, Pp, Pi, Px)
    assert npz >= 0
    return dict(n=n, m=m, Ap=Ap, Ai=Ai, Ax=Ax, l=l, u=u, Pp=Pp, Pi=Pi[:npz].copy(), Px=Px[:npz].copy())


def osqp_solve(p: CsdoParams, n, m, Pp, Pi, Px, q, Ap, Ai, Ax, l, u, x_warm=None,
               max_iter=4000, linsys=0):
    x = np.zeros(n)
    y = np.zeros(max(m, 1))
    st, it, nf, obj = C.c_int(), C.c_int(), C.c_int(), (C.c_double * 4)()
    a = lambda v, dt: np.ascontiguousarray(v, dt)
    xw = a(x_warm, np.float64) if x_warm is not None else None
    lib().orc_osqp_solve(n, m, a(Pp, np.int32), a(Pi, np.int32), a(Px, np.float64), a(q, np.float64),
                         a(Ap, np.int32), a(Ai, np.int32), a(Ax, np.float64), a(l, np.float64),
                         a(u, np.float64), _vp(xw), C.byref(p), max_iter, linsys, x, y.ctypes.data,
                         C.byref(st), C.byref(it), C.byref(nf), obj)
    return dict(x=x, y=y[:m], status=st.value, iters=it.value, n_factor=nf.value, obj=obj[0],
                pri_res=obj[1], dua_res=obj[2], rho=obj[3])


def refine(p: CsdoParams, batch: Batch, linsys: int = 0, nthreads: int = 0):
    """SolverDSQP::SolverDSQP over a batch on the CPU -> (RefineResult, flops)."""
    batch.validate()
    res = RefineResult.allocate(batch)
    cb, cr = batch.to_ctypes(), res.to_ctypes()
    fl = C.c_double()
    rc = lib().orc_refine(C.byref(p), C.byref(cb), C.byref(cr), linsys, nthreads, C.byref(fl))
    if rc != 0:
        raise RuntimeError(f"orc_refine failed: {rc}")
    return res, fl.value


def interpolate_guess(states, actions, goal, n_interp, dt, r, LF, LB, nt_out):
    states = np.ascontiguousarray(states, np.float64)
    out = np.zeros(6 * nt_out)
    g = np.ascontiguousarray(goal, np.float64) if goal is not None else None
    ns = lib().orc_interpolate_guess(states.shape[0], states.reshape(-1),
                                     np.ascontiguousarray(actions, np.int32), _vp(g), n_interp, dt,
                                     r, LF, LB, nt_out, out)
    assert ns > 0
    return out.reshape(6, nt_out), ns
