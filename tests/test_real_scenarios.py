"""Real benchmark geometry: tests/golden/real_scenarios.npz holds 579 of the
reference's own benchmark scenarios (map50by50 agents 5..25 empty/obstacle, room agents 10/20, map100by100
agents25: real map sizes, obstacle lists, starts, goals) together with coarse plans from the stand-in
prioritized planner (tools/coarse_planner.cpp with up to 3 rounds of priority reshuffling; the reference's PBS +
Hybrid A* cannot be built offline).
CPU: the fixture against the YAML files (where /root/reference exists), the YAML quirks, the host chain.
GPU: the whole set as ONE batch -- success rate, collision verdict, and a sample against the oracle."""
import glob
import os

import numpy as np
import pytest

from csdotrajectoryplanning_b200 import pack_instances
from csdotrajectoryplanning_b200 import verdict as V
from csdotrajectoryplanning_b200.driver import instances_from_coarse_plans, run_mapset
from csdotrajectoryplanning_b200.scenario import load_scenario_yaml

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = os.path.join(HERE, "golden", "real_scenarios.npz")
REF = "/root/reference/benchmark"
need_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="needs /root/reference")


@need_ref
def test_fixture_matches_the_yaml_files():
    d = np.load(FIX)
    for i in (0, 57, 200, 333, len(d["name"]) - 1):
        dx, dy, obs, st, gl = load_scenario_yaml(os.path.join(REF, str(d["name"][i])))
        a0, a1, o0, o1 = d["agent_ptr"][i], d["agent_ptr"][i + 1], d["obs_ptr"][i], d["obs_ptr"][i + 1]
        assert (dx, dy) == tuple(d["dims"][i]) and np.array_equal(obs, d["obs"][o0:o1])
        assert np.array_equal(st, d["starts"][a0:a1]) and np.array_equal(gl, d["goals"][a0:a1])
        # every coarse path starts at the scenario's start and ends at its goal (position)
        for a in range(a0, a1):
            s0, s1 = d["st_ptr"][a], d["st_ptr"][a + 1]
            assert np.array_equal(d["states"][s0], st[a - a0]) and np.allclose(d["states"][s1 - 1, :2], gl[a - a0, :2])


@need_ref
def test_yaml_quirks_of_the_benchmark_set():
    """Instance.cc:36-47: a 2-element obstacle gets obsRadius 0.8; null obstacle lists are empty; the dummy
    obstacles [-1,-1] and [-1,-1,0.1] of the agents100/agents70 'empty' folders load as ordinary obstacles."""
    f = sorted(glob.glob(os.path.join(REF, "map100by100/agents100/obstacle/*.yaml")))[0]
    _, _, obs, st, _ = load_scenario_yaml(f)
    assert obs.shape == (50, 3) and np.all(obs[:, 2] == 0.8) and st.shape == (100, 3)
    f = sorted(glob.glob(os.path.join(REF, "map100by100/agents100/empty/*.yaml")))[0]
    _, _, obs, _, _ = load_scenario_yaml(f)
    assert obs.shape == (1, 3) and tuple(obs[0]) == (-1.0, -1.0, 0.8)
    radii = set()
    for f in sorted(glob.glob(os.path.join(REF, "map100by100/agents70/empty/*.yaml"))):
        _, _, obs, _, _ = load_scenario_yaml(f)
        radii |= {tuple(o) for o in obs}
    assert (-1.0, -1.0, 0.1) in radii
    _, _, obs, _, _ = load_scenario_yaml(sorted(glob.glob(os.path.join(REF, "map50by50/agents5/empty/*.yaml")))[0])
    assert obs.shape == (0, 3)
    mx = max(load_scenario_yaml(f)[2].shape[0] for f in glob.glob(os.path.join(REF, "room/agents30/*.yaml")))
    assert mx == 298


def test_host_chain_on_real_geometry(oracle, params):
    """fixture -> InterpolateInitalGuess -> planes -> oracle refine -> verdict on two small real instances."""
    d = np.load(FIX)
    small = [i for i in range(len(d["name"])) if d["agent_ptr"][i + 1] - d["agent_ptr"][i] == 5][:2]
    inst = instances_from_coarse_plans(FIX, params, small)
    for ins in inst:
        assert ins.guess.shape[1] == 6 and ins.guess.shape[2] % 3 == 1          # 3 sub-steps per coarse action
        assert np.abs(ins.guess[:, 4]).max() < 1.3 and np.abs(ins.guess[:, 3]).max() <= np.arctan(1 / 3) + 1e-6
        ins.plane_t, ins.plane_abc, _ = oracle.instance_planes(params, ins.guess)
    b = pack_instances(inst)
    r, _ = oracle.refine(params, b, linsys=0, nthreads=2)
    assert np.all(np.abs(r.inst_status) <= 2)
    for i in range(b.n_inst):
        a0, a1 = int(b.inst_agent_ptr[i]), int(b.inst_agent_ptr[i + 1])
        obs = b.obs[3 * b.obs_ptr[i]:3 * b.obs_ptr[i + 1]].reshape(-1, 3)
        inter, static = V.verdict([V.rounded_solution(r.agent_traj(b, a)) for a in range(a0, a1)], obs)
        assert not inter and not static


@pytest.mark.gpu
def test_gpu_dummy_obstacles_match_oracle(oracle, params, solver):
    """The 'empty' folders of agents100 / agents70 carry one obstacle outside the map ([-1,-1,0.8] / [-1,-1,0.1])."""
    from csdotrajectoryplanning_b200.scenario import synthetic_instance
    inst = []
    for s, dummy in ((301, [-1.0, -1.0, 0.8]), (302, [-1.0, -1.0, 0.1])):
        ins = synthetic_instance(s, 50.0, 5, 0, (8, 14), params)
        ins.obstacles = np.asarray([dummy])
        ins.plane_t, ins.plane_abc, _ = oracle.instance_planes(params, ins.guess)
        inst.append(ins)
    b = pack_instances(inst)
    ro, _ = oracle.refine(params, b, linsys=0, nthreads=2)
    rg = solver.refine(b)
    for k in ("status", "sqp_iters", "admm_iters", "n_factor", "inst_status", "inst_static_legal"):
        assert np.array_equal(getattr(ro, k), getattr(rg, k)), k
    # (the boxes are anchored at the iterate's disc centres, so they inherit its 1e-9-level solver rounding)
    assert np.abs(ro.traj - rg.traj).max() < 1e-6 and np.abs(ro.corridors - rg.corridors).max() < 1e-6


@pytest.mark.gpu
def test_gpu_real_benchmark_set_one_batch(oracle, params, solver, tmp_path):
    """All 579 routed real scenarios (8640 agents) as one batch: success rule of analysis_result.py, collision
    verdict of collision_detection.py on the 3-decimal output, and the oracle on a sample of the same batch."""
    inst = instances_from_coarse_plans(FIX, params)
    n = len(inst)
    assert n == 579 and sum(i.n_agents for i in inst) == 8640
    rep = run_mapset(inst, solver, out_dir=str(tmp_path / "out"), check_collisions=True)
    assert len(rep.files) == n and os.path.getsize(rep.files[0]) > 1000
    s = rep.summary()
    assert s["success_rate"] >= 0.95, s
    assert s["collision_free"] >= 0.9, s
    # oracle on every 12th instance: identical statuses / verdicts, trajectories close on the bulk
    sel = list(range(0, n, 12))
    sub = [inst[i] for i in sel]
    for ins in sub:
        ins.plane_t, ins.plane_abc, _ = oracle.instance_planes(params, ins.guess)
    ob = pack_instances(sub)
    ro, _ = oracle.refine(params, ob, linsys=0, nthreads=os.cpu_count() or 1)
    assert np.array_equal(np.abs(ro.inst_status) <= 2, rep.success[sel])
    agree = float(np.mean(ro.inst_status == rep.solver_status[sel]))
    assert agree >= 0.9, agree
