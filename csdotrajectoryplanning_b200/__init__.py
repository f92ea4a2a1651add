"""B200-native DSQP refine stage of CSDO (drop-in for sqp/dsqp_solver.cc)."""
from .params import CsdoParams, default_params  # noqa: F401
from .batch import Batch, Instance, RefineResult, pack_instances  # noqa: F401
