"""ctypes wrapper of tools/coarse_planner.cpp ("f1-lite"): a deterministic prioritized planner over the
reference's seven motion primitives on the REAL benchmark geometry.  It stands in for the reference's PBS +
Hybrid A* front end (host search, out of scope, needs OMPL) so that real scenarios can be pushed through
InterpolateInitalGuess -> planes -> DSQP; it is not a restatement of PBS.  Test / measurement infrastructure.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libcsdo_planner.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)
        L = C.CDLL(_LIB)
        L.plan_prioritized.argtypes = [C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p]
        L.plan_prioritized.restype = C.c_int
        _lib = L
    return _lib


def plan(dimx: float, dimy: float, obstacles: np.ndarray, starts: np.ndarray, goals: np.ndarray, params,
         max_states: int = 120, max_expansions: int = 400000) -> Tuple[Optional[List], int]:
    """-> (paths, n_failed): paths[a] = (states (n, 3), actions (n - 1,)) ready for InterpolateInitalGuess, or
    None for an agent the planner could not route (n_failed counts them)."""
    obs = np.ascontiguousarray(obstacles, np.float64).reshape(-1, 3)
    st = np.ascontiguousarray(starts, np.float64).reshape(-1, 3)
    gl = np.ascontiguousarray(goals, np.float64).reshape(-1, 3)
    na = st.shape[0]
    ns = np.zeros(na, np.int32)
    states = np.zeros((na, max_states, 3))
    actions = np.zeros((na, max_states), np.int32)
    failed = lib().plan_prioritized(dimx, dimy, obs.shape[0], obs.ctypes.data if obs.size else None, na, st.ctypes.data,
                                    gl.ctypes.data, params.f2x, params.r2x, params.rv, max_states, max_expansions,
                                    ns.ctypes.data, states.ctypes.data, actions.ctypes.data)
    paths = [(states[a, :ns[a]].copy(), actions[a, :ns[a] - 1].copy()) if ns[a] > 0 else None for a in range(na)]
    return paths, int(failed)
