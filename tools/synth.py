"""ctypes wrapper of tools/synth_workload.cpp: seeded synthetic instances, generated in C++.

Measurement infrastructure for bench.py and the full-size tests (BASELINE configs[4] needs 4096 x 100
agents; the numpy generator in csdotrajectoryplanning_b200/scenario.py is kept for the small test cases).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libcsdo_synth.so")
_lib = None


def build() -> None:
    subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = C.CDLL(_LIB)
        L.synth_instance.argtypes = [C.c_uint64, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                     C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                     C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.POINTER(C.c_int)]
        L.synth_instance.restype = C.c_int
        L.interp_paths.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                   C.c_double, C.c_int, C.c_void_p]
        L.interp_paths.restype = C.c_int
        _lib = L
    return _lib


def synth_instance(seed: int, size: float, n_agents: int, n_obs: int, n_actions: Tuple[int, int], params,
                   name: str = "", obs_radius: float = 0.8):
    """Same contract as scenario.synthetic_instance (different generator, so different instances)."""
    from csdotrajectoryplanning_b200.batch import Instance
    L = lib()
    nt_cap = 3 * n_actions[1] + 1
    guess = np.zeros((n_agents, 6, nt_cap))
    obs = np.zeros((max(n_obs, 1), 3))
    nt, no = C.c_int(0), C.c_int(0)   # obs_radius < 0: room-like wall layout with discs of radius -obs_radius
    rc = L.synth_instance(seed, size, n_agents, n_obs, n_actions[0], n_actions[1], obs_radius, params.f2x,
                          params.r2x, params.rv, params.dt, params.LF, params.LB, guess.ctypes.data, nt_cap,
                          C.byref(nt), obs.ctypes.data, C.byref(no))
    if rc:
        raise RuntimeError(f"synth_instance failed ({rc}); lower the density")
    g = guess.reshape(-1)[: n_agents * 6 * nt.value].reshape(n_agents, 6, nt.value).copy()
    return Instance(g, size, size, obs[: no.value].copy(), None, None, name or f"synth_{seed}")


def synth_jobs(jobs, params, threads: Optional[int] = None) -> List:
    """jobs: (seed, size, n_agents, n_obs, (min, max) actions, name, obs_radius)"""
    lib()
    with ThreadPoolExecutor(threads or min(32, os.cpu_count() or 1)) as ex:   # ctypes releases the GIL
        return list(ex.map(lambda a: synth_instance(a[0], a[1], a[2], a[3], a[4], params, a[5], a[6]), jobs))


def synth_batch(shapes: Sequence[Tuple[float, int, int, Tuple[int, int]]], per_shape: int, seed: int, params,
                threads: Optional[int] = None) -> List:
    """shapes: (map size, n_agents, n_obstacles, (min, max) coarse actions); per_shape instances of each."""
    jobs = []
    k = 0
    for (size, na, no, nact) in shapes:
        for j in range(per_shape):
            jobs.append((seed + k, size, na, no, nact, f"map{int(size)}_a{na}_o{no}_n{nact[1]}_ex{j}", 0.8))
            k += 1
    return synth_jobs(jobs, params, threads)


def workload_jobs(name: str, total: int, seed: int = 1234):
    """The job list of a named bench workload; instance i is the same whatever the rank count.
    c5          BASELINE configs[4]: 100x100, 100 agents, 50 obstacles, horizons 127 / 190 / 256 in turn
    map100_a100 BASELINE configs[2]: benchmark/map100by100/agents100/obstacle shape (60 instances there)
    room        BASELINE configs[3]: benchmark/room shape: 100x100, agents 10..50, 130..298 wall discs r = 0.5
    map50       BASELINE configs[1]: the map50by50 sweep, agents 5..25 x {empty, 25 obstacles}
    """
    jobs = []
    for i in range(total):
        if name == "c5":
            size, na, no, nact, rad = C5_SHAPES[i % 3] + (0.8,)
        elif name == "map100_a100":
            size, na, no, nact, rad = 100.0, 100, 50, (40, 58), 0.8
        elif name == "room":
            size, na, no, nact, rad = 100.0, (10, 20, 30, 40, 50)[i % 5], 130 + (37 * i) % 169, (36, 55), -0.5
        elif name == "map50":
            size, na, no, nact, rad = 50.0, (5, 10, 15, 20, 25)[i % 5], (0, 25)[(i // 5) % 2], (12, 30), 0.8
        else:
            raise ValueError(name)
        jobs.append((seed + i, size, na, no, nact, f"{name}_{i}", rad))
    return jobs


# BASELINE.json configs[4]: 100x100 maps, 100 agents (the largest benchmark agent count), 50 obstacles,
# horizons 127 / 190 / 256 (Nt in {128, 192, 256}; a horizon is 3 x coarse actions + 1)
C5_SHAPES = [(100.0, 100, 50, (28, 42)), (100.0, 100, 50, (43, 63)), (100.0, 100, 50, (57, 85))]


def interpolate_paths(n_states: np.ndarray, states: np.ndarray, actions: np.ndarray, goals, params) -> np.ndarray:
    """include/csdo/initial_guess.h::InterpolateInitalGuess on flat arrays (bit-identical to
    scenario.interpolate_initial_guess, which tests/test_initial_guess_cpp.py and tests/test_ref_pins.py pin)."""
    ns = np.ascontiguousarray(n_states, np.int32)
    st = np.ascontiguousarray(states, np.float64)
    ac = np.ascontiguousarray(actions, np.int8)
    gl = np.ascontiguousarray(goals, np.float64) if goals is not None else None
    cap = int(3 * (ns.max() - 1) + 1)
    out = np.zeros((ns.shape[0], 6, cap))
    nt = lib().interp_paths(ns.shape[0], ns.ctypes.data, st.ctypes.data, ac.ctypes.data,
                            gl.ctypes.data if gl is not None else None, params.dt, params.LF, params.LB, cap, out.ctypes.data)
    assert nt > 0
    return out[:, :, :nt].copy()
