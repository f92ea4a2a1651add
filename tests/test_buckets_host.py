"""Host logic of the horizon buckets (csdo_plan_horizon_buckets: no device needed)."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from csdotrajectoryplanning_b200 import binding


def plan(nt, min_count):
    nt = np.ascontiguousarray(nt, np.int32)
    order = np.full(nt.shape[0], -1, np.int32)
    bnt, bcnt = np.zeros(8, np.int32), np.zeros(8, np.int32)
    nb = binding.lib().csdo_plan_horizon_buckets(nt.shape[0], nt.ctypes.data, min_count, order.ctypes.data,
                                                  bnt.ctypes.data, bcnt.ctypes.data, 8)
    assert nb >= 0
    return order, bnt[:nb], bcnt[:nb]


def test_known_plans():
    # a real-map-set-like mix: many short agents, a few long ones in several classes
    nt = [50] * 1500 + [90] * 2500 + [120] * 300 + [150] * 250 + [200] * 200
    order, bnt, bcnt = plan(nt, 2368)
    assert bnt.tolist() == [200, 90] and bcnt.tolist() == [750, 4000]        # one launch per solver family
    order, bnt, bcnt = plan(nt, 200)
    assert bnt.tolist() == [200, 150, 120, 90, 50] and bcnt.tolist() == [200, 250, 300, 2500, 1500]
    order, bnt, bcnt = plan([50] * 3000 + [90] * 2500, 2368)                  # both classes are big enough
    assert bnt.tolist() == [90, 50]
    # a class never moves across the solver boundary, however small it is
    order, bnt, bcnt = plan([40] * 3 + [300] * 5000, 2368)
    assert bnt.tolist() == [300, 40] and bcnt.tolist() == [5000, 3]
    # configs[4] shape at the bench size: three classes that all keep their own launch
    order, bnt, bcnt = plan([127] * 34200 + [190] * 34100 + [256] * 34100, 2368)
    assert bnt.tolist() == [256, 190, 127]
    assert plan([], 10)[1].size == 0
    assert binding.lib().csdo_plan_horizon_buckets(1, np.array([600], np.int32).ctypes.data, 1, np.zeros(1, np.int32).ctypes.data,
                                                   np.zeros(8, np.int32).ctypes.data, np.zeros(8, np.int32).ctypes.data, 8) == -1


@settings(max_examples=200, deadline=None)
@given(st.lists(st.tuples(st.integers(3, 512), st.integers(1, 40)), min_size=1, max_size=30), st.integers(1, 300))
def test_plan_properties(groups, min_count):
    nt = np.concatenate([np.full(c, h, np.int32) for h, c in groups])
    order, bnt, bcnt = plan(nt, min_count)
    assert sorted(order.tolist()) == list(range(nt.shape[0]))                 # a permutation of the agents
    assert 1 <= bnt.shape[0] <= 8 and int(bcnt.sum()) == nt.shape[0]
    assert np.all(np.diff(bnt) < 0)                                          # longest class first
    pos = 0
    for b in range(bnt.shape[0]):
        seg = nt[order[pos:pos + bcnt[b]]]
        pos += bcnt[b]
        assert seg.max() == bnt[b]                                           # the launch is shaped for its longest agent
        assert np.all(np.diff(seg) <= 0)                                     # longest first inside the bucket
        assert (seg <= 96).all() or (seg > 96).all()                         # one solver family per launch
        if b + 1 < bnt.shape[0]:                                             # no agent could have gone to a later (shorter) bucket's class
            assert seg.min() > bnt[b + 1]
