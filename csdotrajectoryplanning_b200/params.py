"""Parameter block of the DSQP refine stage (ctypes mirror of ``csdo_params``).

Values follow the reference's shipped ``config.yaml`` as it is read by
``readAgentConfig`` (common/motion_planning.cc:54-109) and
``readQpSolverConfig`` (sqp/utils.cc:34-59); the OSQP block is what
``osqp_set_default_settings`` of OSQP 0.6.x leaves in force at
sqp/dsqp_solver.cc:480-487.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np


class CsdoParams(C.Structure):
    _fields_ = [
        ("f2x", C.c_double), ("r2x", C.c_double), ("rv", C.c_double),
        ("WB", C.c_double), ("steer_max", C.c_double),
        ("LF", C.c_double), ("LB", C.c_double), ("car_width", C.c_double),
        ("r_trust", C.c_double), ("max_omega", C.c_double), ("max_v", C.c_double),
        ("delta_solution_threshold", C.c_double), ("dt", C.c_double),
        ("max_iter", C.c_int32), ("osqp_max_iter", C.c_int32),
        ("fixed_corridor", C.c_int32),
        ("adaptive_rho_interval", C.c_int32), ("scaling", C.c_int32),
        ("check_termination", C.c_int32), ("adaptive_rho", C.c_int32),
        ("rho", C.c_double), ("sigma", C.c_double), ("alpha", C.c_double),
        ("eps_abs", C.c_double), ("eps_rel", C.c_double),
        ("eps_prim_inf", C.c_double), ("eps_dual_inf", C.c_double),
        ("adaptive_rho_tolerance", C.c_double),
        ("box_ds", C.c_double), ("box_limit", C.c_double),
    ]

    def copy(self) -> "CsdoParams":
        out = CsdoParams()
        C.memmove(C.byref(out), C.byref(self), C.sizeof(CsdoParams))
        return out

    def as_dict(self) -> dict:
        return {name: getattr(self, name) for name, _ in self._fields_}


def default_params(
    r: float = 3.0, deltat: float = 0.706, LF: float = 2.0, LB: float = 1.0,
    car_width: float = 2.0, WB: float = 1.0, max_v: float = 1.0,
    max_omega: float = 0.07, decelerate_factor: float = 0.8,
    num_interpolation: int = 2, r_trust: float = 2.0, max_iter: int = 10,
    delta_solution_threshold: float = 1.0, osqp_max_iter: int = 400,
    fixed_corridor: bool = False, adaptive_rho_interval: int = 25,
) -> CsdoParams:
    """config.yaml -> parameters, with the reference's float/double mix.

    ``Constants`` members are ``float`` (common/motion_planning.h:9-49), so
    f2x/r2x/rv are rounded to float and ``dt`` starts from a float product
    (sqp/utils.cc:55-56).
    """
    f32 = np.float32
    p = CsdoParams()
    LFf, LBf, Wf, rf, dtf = f32(LF), f32(LB), f32(car_width), f32(r), f32(deltat)
    # motion_planning.cc:82-85 (double expression stored to float)
    p.f2x = float(f32(1 / 4.0 * (3.0 * float(LFf) - float(LBf))))
    p.r2x = float(f32(1 / 4.0 * (float(LFf) - 3.0 * float(LBf))))
    p.rv = float(f32(1.0 / 2.0 * math.pow(
        math.pow(float(f32(LFf + LBf)), 2) / 4 + float(f32(Wf * Wf)), 0.5)))
    p.WB = float(f32(WB))
    p.steer_max = math.atan(p.WB / float(rf))  # dsqp_solver.cc:1178
    p.LF, p.LB, p.car_width = float(LFf), float(LBf), float(Wf)
    p.r_trust = r_trust
    p.max_omega = max_omega
    p.max_v = max_v
    p.delta_solution_threshold = delta_solution_threshold
    # utils.cc:55-56: float product, then double divisions left to right
    p.dt = float(f32(rf * dtf)) / max_v / (num_interpolation + 1) / decelerate_factor
    p.max_iter = int(max_iter)
    p.osqp_max_iter = int(osqp_max_iter)
    p.fixed_corridor = int(bool(fixed_corridor))
    p.adaptive_rho_interval = int(adaptive_rho_interval)
    p.scaling = 10
    p.check_termination = 25
    p.adaptive_rho = 1
    p.rho, p.sigma, p.alpha = 0.1, 1e-6, 1.6
    p.eps_abs = p.eps_rel = 1e-3
    p.eps_prim_inf = p.eps_dual_inf = 1e-4
    p.adaptive_rho_tolerance = 5.0
    p.box_ds, p.box_limit = 0.1, 10.0
    return p
