"""ctypes binding of libcsdo_dsqp.so (include/csdo_dsqp.h).

The library is built in-tree by ``__graft_entry__.build()`` /
``make -C csdotrajectoryplanning_b200/csrc``.  There is no CPU fallback: if the
library is missing this module raises, and every compute call fails with
CSDO_ERR_CUDA when no B200 is visible.
"""
from __future__ import annotations

import ctypes as C
import os

from .batch import CsdoBatch, CsdoLaunchInfo, CsdoResult
from .params import CsdoParams

_HERE = os.path.dirname(os.path.abspath(__file__))
# CSDO_LIB (developer knob): another build of the same library, e.g. one with -DCSDO_DEV_TIMERS
LIB_PATH = os.environ.get("CSDO_LIB") or os.path.join(_HERE, "csrc", "libcsdo_dsqp.so")

CSDO_OK, CSDO_ERR_INVALID, CSDO_ERR_CUDA, CSDO_ERR_UNSUPPORTED, CSDO_ERR_NOMEM = 0, 1, 2, 3, 4

EXPORTS = [
    "csdo_default_params", "csdo_version", "csdo_create", "csdo_destroy", "csdo_last_error",
    "csdo_refine", "csdo_refine_device", "csdo_refine_device_hinted", "csdo_plan_horizon_buckets", "csdo_last_launch", "csdo_corridors",
    "csdo_planes_count", "csdo_planes_fill", "csdo_planes_fill_partners", "csdo_planes_from_pairs",
    "csdo_planes_count_device", "csdo_planes_fill_device", "csdo_sync", "csdo_aggregate_status_device", "csdo_measure_fp64_peak",
]

_lib = None


class CsdoError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"csdo error {code}: {msg}")
        self.code = code


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(the DSQP refine path has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        H = C.c_void_p
        L.csdo_default_params.argtypes = [C.POINTER(CsdoParams)]
        L.csdo_default_params.restype = None
        L.csdo_version.restype = C.c_char_p
        L.csdo_create.argtypes = [C.POINTER(CsdoParams), C.c_int, C.POINTER(H)]
        L.csdo_create.restype = C.c_int
        L.csdo_destroy.argtypes = [H]
        L.csdo_destroy.restype = None
        L.csdo_last_error.argtypes = [H]
        L.csdo_last_error.restype = C.c_char_p
        L.csdo_refine.argtypes = [H, C.POINTER(CsdoBatch), C.POINTER(CsdoResult)]
        L.csdo_refine.restype = C.c_int
        L.csdo_refine_device.argtypes = [H, C.POINTER(CsdoBatch), C.POINTER(CsdoResult), C.c_int,
                                         C.c_int, C.c_void_p]
        L.csdo_refine_device.restype = C.c_int
        L.csdo_refine_device_hinted.argtypes = [H, C.POINTER(CsdoBatch), C.POINTER(CsdoResult), C.c_int, C.c_void_p,
                                                C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.csdo_refine_device_hinted.restype = C.c_int
        L.csdo_plan_horizon_buckets.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.csdo_plan_horizon_buckets.restype = C.c_int
        L.csdo_last_launch.argtypes = [H, C.POINTER(CsdoLaunchInfo)]
        L.csdo_last_launch.restype = C.c_int
        L.csdo_corridors.argtypes = [H, C.POINTER(CsdoBatch), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.csdo_corridors.restype = C.c_int
        L.csdo_planes_count.argtypes = [H, C.POINTER(CsdoBatch), C.c_void_p, C.c_void_p]
        L.csdo_planes_count.restype = C.c_int
        L.csdo_planes_fill.argtypes = [H, C.POINTER(CsdoBatch), C.c_void_p, C.c_void_p, C.c_void_p]
        L.csdo_planes_fill.restype = C.c_int
        L.csdo_planes_fill_partners.argtypes = [H, C.POINTER(CsdoBatch), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.csdo_planes_fill_partners.restype = C.c_int
        L.csdo_planes_from_pairs.argtypes = [H, C.POINTER(CsdoBatch), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p]
        L.csdo_planes_from_pairs.restype = C.c_int
        L.csdo_planes_count_device.argtypes = [H, C.POINTER(CsdoBatch), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.POINTER(C.c_int64), C.c_void_p]
        L.csdo_planes_count_device.restype = C.c_int
        L.csdo_planes_fill_device.argtypes = [H, C.POINTER(CsdoBatch), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p]
        L.csdo_planes_fill_device.restype = C.c_int
        L.csdo_aggregate_status_device.argtypes = [H, C.POINTER(CsdoBatch), C.POINTER(CsdoResult), C.c_void_p]
        L.csdo_aggregate_status_device.restype = C.c_int
        L.csdo_sync.argtypes = [H]
        L.csdo_sync.restype = C.c_int
        L.csdo_measure_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
        L.csdo_measure_fp64_peak.restype = C.c_int
        _lib = L
    return _lib
