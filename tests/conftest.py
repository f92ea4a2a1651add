import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def params():
    from csdotrajectoryplanning_b200 import default_params
    return default_params()


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


def make_batch(oracle, params, seeds, size=50.0, na=5, no=12, acts=(8, 16)):
    """Small synthetic batch with planes built by the oracle (test infrastructure)."""
    from csdotrajectoryplanning_b200 import pack_instances
    from csdotrajectoryplanning_b200.scenario import synthetic_instance
    inst = []
    for s in seeds:
        ins = synthetic_instance(s, size, na, no, acts, params)
        ins.plane_t, ins.plane_abc, _ = oracle.instance_planes(params, ins.guess)
        inst.append(ins)
    return pack_instances(inst)


@pytest.fixture(scope="session")
def small_batch(oracle, params):
    return make_batch(oracle, params, [11, 12, 13])


@pytest.fixture(scope="session")
def solver(params):
    from csdotrajectoryplanning_b200.solver import DsqpSolver
    s = DsqpSolver(params)
    yield s
    s.close()
