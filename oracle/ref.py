"""TEST INFRASTRUCTURE -- ctypes loader of oracle/_ref/libcsdo_ref.so: the REFERENCE'S OWN
sqp/corridor.cc and sqp/inter_agent_cons.cc compiled unmodified (recipe: `make -C oracle ref`, see
oracle/Makefile).  It pins the restatement in oracle/dsqp_restate.c and mints the fixtures of
tests/golden/ref_pins.npz.  The library only exists where /root/reference does (or where a prebuilt
copy travelled with the snapshot); `available()` says which.  Never imported by the product path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libcsdo_ref.so")
_lib = None
_dp = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build() -> None:
    subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def available() -> bool:
    if not os.path.exists(_SO) and os.path.isdir("/root/reference/sqp"):
        try:
            build()
        except Exception:
            return False
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise ImportError("oracle/_ref/libcsdo_ref.so is not built (needs /root/reference)")
        L = C.CDLL(_SO)
        L.ref_obstacle_order.argtypes = [C.c_void_p, C.c_int, _ip]
        L.ref_obstacle_order.restype = C.c_int
        L.ref_generate_box.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_int, _dp, _ip]
        L.ref_generate_box.restype = None
        L.ref_calc_corridors.argtypes = [_dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_int, _dp]
        L.ref_calc_corridors.restype = C.c_int
        L.ref_instance_planes.argtypes = [_dp, C.c_int, C.c_int, C.c_double, _ip, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.POINTER(C.c_int)]
        L.ref_instance_planes.restype = C.c_int
        L.ref_interpolate_guess.argtypes = [C.c_int, _ip, _dp, _ip, C.c_void_p, C.c_int, C.c_double, C.c_int, _dp]
        L.ref_interpolate_guess.restype = C.c_int
        L.ref_dump_solutions.argtypes = [C.c_char_p, _dp, C.c_int, C.c_int, _dp]
        L.ref_dump_solutions.restype = None
        _lib = L
    return _lib


def _obs(obs):
    o = np.ascontiguousarray(obs, np.float64).reshape(-1, 3)
    return o, (o.ctypes.data if o.size else None)


def obstacle_order(obs) -> np.ndarray:
    """Iteration order of the reference's std::unordered_set<Location> filled in the given order."""
    o, ptr = _obs(obs)
    out = np.zeros(max(o.shape[0], 1), np.int32)
    n = lib().ref_obstacle_order(ptr, o.shape[0], out)
    return out[:n].copy()


def generate_box(dimx, dimy, x, y, obs):
    o, ptr = _obs(obs)
    box, st = np.zeros(4), np.zeros(2, np.int32)
    lib().ref_generate_box(dimx, dimy, x, y, ptr, o.shape[0], box, st)
    return box, st


def calc_corridors(guess, dimx, dimy, obs):
    """guess (Na,6,Nt) -> (corridors (Na,8,Nt), initial_static_legal)."""
    g = np.ascontiguousarray(guess, np.float64)
    na, _, nt = g.shape
    o, ptr = _obs(obs)
    corr = np.zeros((na, 8, nt))
    ok = lib().ref_calc_corridors(g.reshape(-1), na, nt, dimx, dimy, ptr, o.shape[0], corr.reshape(-1))
    return corr, bool(ok)


def instance_planes(guess, r_trust=2.0):
    g = np.ascontiguousarray(guess, np.float64)
    na, _, nt = g.shape
    cnt = np.zeros(na, np.int32)
    npairs = C.c_int()
    lib().ref_instance_planes(g.reshape(-1), na, nt, r_trust, cnt, None, None, None, C.byref(npairs))
    ptr = np.zeros(na + 1, np.int32)
    ptr[1:] = np.cumsum(cnt)
    pt = np.zeros(max(int(ptr[-1]), 1), np.int32)
    pabc = np.zeros(max(int(ptr[-1]), 1) * 12)
    legal = lib().ref_instance_planes(g.reshape(-1), na, nt, r_trust, cnt, pt.ctypes.data, pabc.ctypes.data,
                                      ptr.ctypes.data, C.byref(npairs))
    pts = [pt[ptr[a]:ptr[a + 1]].copy() for a in range(na)]
    pabcs = [pabc[12 * ptr[a]:12 * ptr[a + 1]].reshape(-1, 12).copy() for a in range(na)]
    return pts, pabcs, bool(legal), npairs.value


def interpolate_guess(paths, goals, num_interpolation, dt):
    """paths: list of (states (n,3), actions (n-1,)) -> guess (Na,6,Nt) via InterpolateInitalGuess."""
    na = len(paths)
    ns = np.asarray([len(p[0]) for p in paths], np.int32)
    st = np.ascontiguousarray(np.concatenate([np.asarray(p[0], np.float64).reshape(-1, 3) for p in paths])).reshape(-1)
    ac = np.ascontiguousarray(np.concatenate([np.asarray(p[1], np.int32).reshape(-1) for p in paths]), np.int32)
    cap = int((ns.max() - 1) * (num_interpolation + 1) + 1)
    out = np.zeros((na, 6, cap))
    gl = np.ascontiguousarray(goals, np.float64).reshape(-1) if goals is not None else None
    nt = lib().ref_interpolate_guess(na, ns, st, ac, gl.ctypes.data if gl is not None else None, num_interpolation,
                                     dt, cap, out.reshape(-1))
    assert nt > 0
    return out[:, :, :nt].copy()


def dump_solutions(sol, stat) -> str:
    """dumpSolutions text for sol (Na,6,Nt) planes and the 10 SolutionStatistics values."""
    s = np.ascontiguousarray(sol, np.float64)
    na, _, nt = s.shape
    with tempfile.NamedTemporaryFile(suffix=".yaml", delete=False) as f:
        name = f.name
    try:
        lib().ref_dump_solutions(name.encode(), s.reshape(-1), na, nt, np.ascontiguousarray(stat, np.float64))
        return open(name).read()
    finally:
        os.unlink(name)
