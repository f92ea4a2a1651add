"""Solution files of the reference: the YAML that `dumpSolutions` writes
(sqp/inter_agent_cons.cc:413-455) and the status rule `scripts/analysis_result.py:53-101` reads back.

Writer and readers only -- no arithmetic of the refine path lives here.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

DEG = 180 / 3.14  # the reference prints steer and omega as value * 180 / 3.14 (inter_agent_cons.cc:447,451)
SOLVER_THRESHOLD = 2  # analysis_result.py: success <=> abs(solver_status) <= 2


@dataclass
class SolutionStatistics:  # sqp/common.h:25-36
    cost: float = -1
    makespan: float = -1
    flowtime: float = -1
    runtime: float = -1
    rt_search: float = -1
    rt_preprocess: float = -1
    rt_optimization: float = -1
    rt_max_optimization: float = -1
    search_status: int = 2
    solver_status: int = 0


def _f(v: float) -> str:
    """std::fixed << std::setprecision(3)"""
    return f"{float(v):.3f}"


def format_solutions(trajs: np.ndarray, stat: SolutionStatistics) -> str:
    """trajs: (Na, 6, Nt) planes x,y,yaw,steer,v,w.  Same text as dumpSolutions (inter_agent_cons.cc:413-455):
    v and omega are omitted at the last step, steer/omega are scaled by 180/3.14."""
    trajs = np.asarray(trajs, np.float64)
    na, _, nt = trajs.shape
    out = ["statistics:",
           f"  cost: {_f(stat.cost)}",
           f"  makespan: {_f(stat.makespan)}",
           f"  flowtime: {_f(stat.flowtime)}",
           f"  runtime: {_f(stat.runtime)}",
           f"  runtime_search: {_f(stat.rt_search)}",
           f"  runtime_preprocess: {_f(stat.rt_preprocess)}",
           f"  runtime_optimization: {_f(stat.rt_optimization)}",
           f"  runtime_decentralized_optimization: {_f(stat.rt_max_optimization)}",
           f"  search_status: {int(stat.search_status)}",
           f"  solver_status: {int(stat.solver_status)}",
           "schedule:"]
    for a in range(na):
        out.append(f"  agent{a}:")
        x, y, yaw, steer, v, w = trajs[a]
        for t in range(nt):
            out.append(f"    - x: {_f(x[t])}")
            out.append(f"      y: {_f(y[t])}")
            out.append(f"      yaw: {_f(yaw[t])}")
            out.append(f"      steer: {_f(steer[t] * 180 / 3.14)}")
            out.append(f"      t: {t}")
            if t == nt - 1:
                continue
            out.append(f"      v: {_f(v[t])}")
            out.append(f"      omega: {_f(w[t] * 180 / 3.14)}")
    return "\n".join(out) + "\n"


def format_corridors(corridors: np.ndarray, guess: np.ndarray, f2x: float, r2x: float) -> str:
    """The file `./csdo --dump_corridor` writes (dumpCorridors, sqp/utils.cc:62-89).  corridors: (Na, 8, Nt)
    planes xf_min, xf_max, yf_min, yf_max, xr_min, xr_max, yr_min, yr_max; guess: (Na, >=3, Nt) planes x, y, yaw of
    x0_bar.  Per agent and step two rows "[disc centre x, y, x_min, x_max, y_min, y_max]" (front, rear) in the
    stream's default format (6 significant digits); the centres pass through State's float members."""
    corridors, guess = np.asarray(corridors, np.float64), np.asarray(guess, np.float64)
    na, _, nt = corridors.shape
    g = lambda v: f"{float(v):.6g}"
    f2, r2 = np.float64(np.float32(f2x)), np.float64(np.float32(r2x))
    out = []
    for a in range(na):
        out.append(f"agent{a}:")
        x, y, yaw = guess[a][0], guess[a][1], guess[a][2]
        c = corridors[a]
        for t in range(nt):
            cs, sn = np.cos(yaw[t]), np.sin(yaw[t])
            xf, yf = np.float32(x[t] + f2 * cs), np.float32(y[t] + f2 * sn)
            xr, yr = np.float32(x[t] + r2 * cs), np.float32(y[t] + r2 * sn)
            out.append(f"  - [{g(xf)}, {g(yf)}, {g(c[0][t])}, {g(c[1][t])}, {g(c[2][t])}, {g(c[3][t])}]")
            out.append(f"  - [{g(xr)}, {g(yr)}, {g(c[4][t])}, {g(c[5][t])}, {g(c[6][t])}, {g(c[7][t])}]")
    return "\n".join(out) + "\n"


def dump_corridors(path: str, corridors: np.ndarray, guess: np.ndarray, f2x: float, r2x: float) -> None:
    with open(path, "w") as f:
        f.write(format_corridors(corridors, guess, f2x, r2x))


def dump_solutions(path: str, trajs: np.ndarray, stat: SolutionStatistics) -> None:
    with open(path, "w") as f:
        f.write(format_solutions(trajs, stat))


def read_solution_status(path: str) -> Tuple[SolutionStatistics, bool]:
    """Header parse of analysis_result.py:53-101 (line positions, not YAML); returns (statistics, success)."""
    with open(path) as f:
        lines = [f.readline() for _ in range(11)]
    val = lambda i: float(lines[i].split()[1])
    st = SolutionStatistics(cost=val(1), makespan=val(2), flowtime=val(3), runtime=val(4), rt_search=val(5),
                            rt_preprocess=val(6), rt_optimization=val(7), rt_max_optimization=val(8),
                            search_status=int(val(9)), solver_status=int(val(10)))
    return st, abs(st.solver_status) <= SOLVER_THRESHOLD


def load_solutions(path: str) -> np.ndarray:
    """schedule -> (Na, 6, Nt) in the units of the solver (steer/omega back to radians via 3.14/180).
    The file carries 3 decimals: this is the lossy hand-off format the reference offers (`_guesses.yaml`
    written by csdo.cc:139 when dump_initial_guess is set), not a bit-exact channel."""
    import yaml
    with open(path) as f:
        doc = yaml.safe_load(f)
    sched = doc["schedule"]
    names = sorted(sched.keys(), key=lambda s: int(s[5:]))
    nt = max(len(sched[n]) for n in names)
    out = np.zeros((len(names), 6, nt))
    for a, n in enumerate(names):
        for t, s in enumerate(sched[n]):
            out[a, 0, t], out[a, 1, t], out[a, 2, t] = s["x"], s["y"], s["yaw"]
            out[a, 3, t] = s["steer"] / DEG
            out[a, 4, t] = s.get("v", 0.0)
            out[a, 5, t] = s.get("omega", 0.0) / DEG
    return out


def rounded_from_file(path: str) -> List[np.ndarray]:
    """Trajectories (x, y, yaw) as visualize.py reads them, one (3, Nt) array per agent."""
    tr = load_solutions(path)
    return [tr[a, :3].copy() for a in range(tr.shape[0])]
