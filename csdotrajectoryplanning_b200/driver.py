"""Batch driver: what scripts/test_through_benchmark.sh does one process per
instance -- run the refine stage on every scenario of a map set and write one solution file each --
done as ONE GPU batch: all instances of the set go through a single csdo_planes_* / csdo_refine call.

Inputs per instance: the benchmark scenario YAML (Instance.cc:6-63: map dimensions, obstacles) and the
interpolated initial guess x0_bar in the reference's own dump format (`<name>_guesses.yaml`, written by
csdo.cc:139 when dump_initial_guess is set).  The PBS / Hybrid-A* front end that produces x0_bar is out of
scope here (row f1); guesses can also be handed over as arrays.
"""
from __future__ import annotations

import glob
import os
import time
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import verdict as V
from .batch import Instance, pack_instances
from .output import SolutionStatistics, dump_corridors, dump_solutions, load_solutions
from .scenario import load_scenario_yaml


@dataclass
class MapsetReport:
    names: List[str]
    solver_status: np.ndarray       # per instance, dsqp_solver.cc:1224-1243
    search_status: np.ndarray       # 2, or 1 when the initial guess is not legal (csdo.cc:121-124,154-156)
    success: np.ndarray             # analysis_result.py: abs(solver_status) <= 2
    collisions: List[int] = field(default_factory=list)  # per instance: inter-vehicle + static, rounded output
    refine_seconds: float = 0.0
    files: List[str] = field(default_factory=list)

    @property
    def success_rate(self) -> float:
        return float(np.mean(self.success)) if len(self.success) else 0.0

    def summary(self) -> Dict[str, float]:
        return {"instances": len(self.names), "success_rate": self.success_rate,
                "collision_free": float(np.mean([c == 0 for c in self.collisions])) if self.collisions else float("nan"),
                "refine_seconds": self.refine_seconds,
                "refine_ms_per_instance": 1e3 * self.refine_seconds / max(1, len(self.names))}


def collect_mapset(scenario_paths: Sequence[str], guess_dir: str) -> List[Instance]:
    """Pairs every scenario YAML with `<stem>_guesses.yaml` (or `<stem>.yaml`) under guess_dir."""
    files: List[str] = []
    for p in scenario_paths:
        files += sorted(glob.glob(os.path.join(p, "**", "*.yaml"), recursive=True)) if os.path.isdir(p) else [p]
    out = []
    for f in files:
        stem = os.path.splitext(os.path.basename(f))[0]
        cand = [os.path.join(guess_dir, stem + "_guesses.yaml"), os.path.join(guess_dir, stem + ".yaml")]
        g = next((c for c in cand if os.path.exists(c)), None)
        if g is None:
            continue
        dimx, dimy, obs, _, _ = load_scenario_yaml(f)
        out.append(Instance(guess=load_solutions(g), dimx=dimx, dimy=dimy, obstacles=obs, name=stem))
    return out


def run_mapset(instances: Sequence[Instance], solver, out_dir: Optional[str] = None,
               check_collisions: bool = True, dump_corridor: bool = False) -> MapsetReport:
    """One batch through planes + refine; optional solution files (dumpSolutions format) and verdicts.
    dump_corridor: also write <name>_corridors.yaml like `./csdo --dump_corridor` (csdo.cc:163-165)."""
    batch = pack_instances(list(instances))
    batch, inter_legal = solver.planes(batch)
    t0 = time.perf_counter()
    res = solver.refine(batch)
    dt = time.perf_counter() - t0
    n = batch.n_inst
    search = np.where((inter_legal != 0) & (res.inst_static_legal != 0), 2, 1).astype(np.int32)
    rep = MapsetReport(names=[i.name for i in instances], solver_status=res.inst_status.copy(),
                       search_status=search, success=np.abs(res.inst_status) <= 2, refine_seconds=dt)
    if out_dir:
        os.makedirs(out_dir, exist_ok=True)
    for i in range(n):
        a0, a1 = int(batch.inst_agent_ptr[i]), int(batch.inst_agent_ptr[i + 1])
        trajs = np.stack([res.agent_traj(batch, a) for a in range(a0, a1)])
        if check_collisions:
            obs = batch.obs[3 * batch.obs_ptr[i]:3 * batch.obs_ptr[i + 1]].reshape(-1, 3)
            inter, static = V.verdict([V.rounded_solution(t) for t in trajs], obs)
            rep.collisions.append(len(inter) + len(static))
        if out_dir:
            stat = SolutionStatistics(rt_optimization=dt, rt_max_optimization=dt / max(1, n),
                                      search_status=int(search[i]), solver_status=int(res.inst_status[i]))
            path = os.path.join(out_dir, (instances[i].name or f"instance{i}") + ".yaml")
            dump_solutions(path, trajs, stat)
            rep.files.append(path)
            if dump_corridor:
                corr = np.stack([res.agent_corridor(batch, a) for a in range(a0, a1)])
                nt = int(batch.inst_nt[i])
                guess = np.stack([batch.guess[6 * batch.agent_off[a]:6 * batch.agent_off[a + 1]].reshape(6, nt)
                                  for a in range(a0, a1)])
                dump_corridors(path[:-5] + "_corridors.yaml", corr, guess, solver.params.f2x, solver.params.r2x)
    return rep


def instances_from_coarse_plans(npz_path: str, params, select: Optional[Sequence[int]] = None) -> List[Instance]:
    """Instances from a coarse-plan fixture (tests/golden/real_scenarios.npz: real benchmark geometry + coarse
    (state, action) paths): InterpolateInitalGuess (scenario.interpolate_initial_guess, goals snapped as in
    inter_agent_cons.cc:149-151) -> x0_bar, real map size and obstacles."""
    from .scenario import interpolate_initial_guess
    d = np.load(npz_path)
    fast = None
    try:      # the C++ header (include/csdo/initial_guess.h, bit-identical) when the tools library is built
        from tools import synth
        fast = synth.interpolate_paths
    except Exception:
        pass
    out = []
    for i in (range(len(d["name"])) if select is None else select):
        a0, a1 = int(d["agent_ptr"][i]), int(d["agent_ptr"][i + 1])
        if fast is not None:
            s0, s1 = int(d["st_ptr"][a0]), int(d["st_ptr"][a1])
            guess = fast(np.diff(d["st_ptr"][a0:a1 + 1]), d["states"][s0:s1], d["actions"][s0:s1], d["goals"][a0:a1], params)
        else:
            paths = []
            for a in range(a0, a1):
                s0, s1 = int(d["st_ptr"][a]), int(d["st_ptr"][a + 1])
                paths.append((d["states"][s0:s1], d["actions"][s0:s1 - 1].astype(np.int32)))
            guess = interpolate_initial_guess(paths, d["goals"][a0:a1], params)
        o0, o1 = int(d["obs_ptr"][i]), int(d["obs_ptr"][i + 1])
        out.append(Instance(guess=guess, dimx=float(d["dims"][i, 0]), dimy=float(d["dims"][i, 1]),
                            obstacles=d["obs"][o0:o1].copy(), name=str(d["name"][i]).replace("/", "__")[:-5]))
    return out
