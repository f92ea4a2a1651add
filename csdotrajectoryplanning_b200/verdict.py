"""Collision verdict and output format of the reference, restated headless.

* :func:`format_solution_value` / :func:`rounded_solution` -- the 3-decimal fixed format of
  ``dumpSolutions`` (sqp/inter_agent_cons.cc:413-455), which is all the verdict scripts ever see.
* :func:`collision_rect_and_rect`, :func:`collision_circle_and_rect` -- scripts/collision_detection.py
  :20-96 (separating axes on rear-axle-anchored LFxLB rectangles, circle vs rectangle).
* :func:`verdict` -- the per-frame loop of scripts/visualize.py:220-249 (framesPerMove = 1: one frame
  per integer time step): all agent pairs, then every agent against every obstacle.
* :func:`success` -- scripts/analysis_result.py:84-87: ``abs(solver_status) <= 2``.
Pinned against the reference's own functions by tests/golden/make_verdict_golden.py.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np

# scripts/visualize.py:20-27
LF, LB, CAR_WIDTH, OBS_RADIUS_VIS = 2.0, 1.0, 2.0, 1.0
SOLVER_THRESHOLD = 2  # scripts/analysis_result.py:10


def rounded_solution(traj: np.ndarray) -> np.ndarray:
    """x, y, yaw as dumpSolutions prints them (std::fixed, setprecision(3)) and a YAML reader parses them."""
    return np.array([[float("%.3f" % v) for v in row] for row in np.asarray(traj)[:3]])


def _rect(p):
    """Rectangle(..., pos='rear_axle_center', lb=LB): centre shifted by length/2 - lb (collision_geometry.py:9-29)."""
    length = LF + LB
    d = length / 2 - LB
    return (p[0] + d * math.cos(p[2]), p[1] + d * math.sin(p[2]), p[2], length, CAR_WIDTH)


def collision_rect_and_rect(r1, r2) -> bool:
    shift_x, shift_y = r2[0] - r1[0], r2[1] - r1[1]
    cos_v, sin_v, cos_o, sin_o = math.cos(r1[2]), math.sin(r1[2]), math.cos(r2[2]), math.sin(r2[2])
    half_l_v, half_w_v, half_l_o, half_w_o = r1[3] / 2, r1[4] / 2, r2[3] / 2, r2[4] / 2
    dx1, dy1, dx2, dy2 = cos_v * r1[3] / 2, sin_v * r1[3] / 2, sin_v * r1[4] / 2, -cos_v * r1[4] / 2
    dx3, dy3, dx4, dy4 = cos_o * r2[3] / 2, sin_o * r2[3] / 2, sin_o * r2[4] / 2, -cos_o * r2[4] / 2
    return ((abs(shift_x * cos_v + shift_y * sin_v) <=
             abs(dx3 * cos_v + dy3 * sin_v) + abs(dx4 * cos_v + dy4 * sin_v) + half_l_v)
            and (abs(shift_x * sin_v - shift_y * cos_v) <=
                 abs(dx3 * sin_v - dy3 * cos_v) + abs(dx4 * sin_v - dy4 * cos_v) + half_w_v)
            and (abs(shift_x * cos_o + shift_y * sin_o) <=
                 abs(dx1 * cos_o + dy1 * sin_o) + abs(dx2 * cos_o + dy2 * sin_o) + half_l_o)
            and (abs(shift_x * sin_o - shift_y * cos_o) <=
                 abs(dx1 * sin_o - dy1 * cos_o) + abs(dx2 * sin_o - dy2 * cos_o) + half_w_o))


def _vertices(r) -> np.ndarray:
    v_d = np.array([[r[3], r[4]], [-r[3], r[4]], [-r[3], -r[4]], [r[3], -r[4]]]) / 2
    c, s = math.cos(r[2]), math.sin(r[2])
    rot = np.array([[c, -s], [s, c]])
    return v_d @ rot.T + np.array([r[0], r[1]])


def collision_circle_and_rect(circle, r) -> bool:
    cx, cy, cr = circle
    vertices = _vertices(r)
    d_min, ind = np.inf, -1
    for i_v in range(vertices.shape[0]):
        d = ((vertices[i_v, 0] - cx) ** 2 + (vertices[i_v, 1] - cy) ** 2) ** 0.5 - cr
        if d < d_min:
            d_min, ind = d, i_v
    if d_min < 0:
        return True
    yaw = r[2]
    axes = [[math.cos(yaw), math.sin(yaw)], [-math.sin(yaw), math.cos(yaw)]]
    ap = vertices[ind, :] - np.array([cx, cy])
    norm = math.sqrt(ap[0] ** 2 + ap[1] ** 2)
    axes.append((ap[0] / norm, ap[1] / norm))
    for axis in axes:
        dots = [v[0] * axis[0] + v[1] * axis[1] for v in vertices]
        pa = [min(dots), max(dots)]
        pc = cx * axis[0] + cy * axis[1]
        pb = [pc - cr, pc + cr]
        if not (min(pa) <= max(pb) and min(pb) <= max(pa)):
            return False
    return True


# Pre-filters of verdict(): they only skip pairs for which the exact tests above provably return False.
# Circumradius of the LF+LB by CAR_WIDTH rectangle; two rectangles whose centres are further apart than 2 R are
# disjoint and the 4-axis SAT is exact for rectangles.  Circle vs rectangle: the third axis of
# collision_circle_and_rect points from the circle centre c to the nearest vertex v*; every vertex w projects to
# (w - c).axis >= |v* - c| - |w - v*| >= (D - R) - 2 R, so that axis separates as soon as D > radius + 3 R.
_R_CIRC = math.hypot((LF + LB) / 2, CAR_WIDTH / 2)


def verdict(trajs: List[np.ndarray], obstacles: np.ndarray) -> Tuple[List[Tuple[int, int, int]], List[Tuple[int, int, int]]]:
    """trajs: per agent (>=3, Nt) arrays (x, y, yaw rows); obstacles (No, 2|3).
    -> (inter collisions [(t, i, j)], static collisions [(t, agent, obstacle index)])."""
    na = len(trajs)
    nt = max(t.shape[1] for t in trajs)
    obs = np.asarray(obstacles, np.float64)
    obs = obs.reshape(-1, obs.shape[-1] if obs.size else 3)
    rad = obs[:, 2] if obs.shape[1] == 3 else np.full(obs.shape[0], OBS_RADIUS_VIS)
    inter, static = [], []
    for f in range(nt):
        pos = [t[:3, min(f, t.shape[1] - 1)] for t in trajs]
        rects = [_rect(p) for p in pos]
        cen = np.array([[r[0], r[1]] for r in rects]).reshape(na, 2)
        d2 = ((cen[:, None, :] - cen[None, :, :]) ** 2).sum(-1)
        close = d2 <= (2 * _R_CIRC + 1e-6) ** 2
        for ai in range(na):
            for aj in np.nonzero(close[ai, ai + 1:])[0] + ai + 1:
                if collision_rect_and_rect(rects[ai], rects[int(aj)]):
                    inter.append((f, ai, int(aj)))
        if obs.shape[0]:
            do2 = ((cen[:, None, :] - obs[None, :, :2]) ** 2).sum(-1)
            near = do2 <= (rad[None, :] + 3 * _R_CIRC + 1e-6) ** 2
            for a in range(na):
                for oi in np.nonzero(near[a])[0]:
                    o = obs[int(oi)]
                    if collision_circle_and_rect((o[0], o[1], rad[int(oi)]), rects[a]):
                        static.append((f, a, int(oi)))
    return inter, static


def success(solver_status: int) -> bool:
    return abs(int(solver_status)) <= SOLVER_THRESHOLD
