"""GPU test of the map-set batch driver: scenario YAMLs + `_guesses.yaml` files in the
reference's formats -> one GPU batch -> solution files; checked against the CPU oracle on the same
(3-decimal) guesses."""
import numpy as np
import pytest
import yaml

from csdotrajectoryplanning_b200 import pack_instances
from csdotrajectoryplanning_b200.driver import collect_mapset, run_mapset
from csdotrajectoryplanning_b200.output import SolutionStatistics, dump_solutions, load_solutions, read_solution_status
from csdotrajectoryplanning_b200.scenario import synthetic_instance

pytestmark = pytest.mark.gpu


def _write_scenario(path, ins):
    doc = {"agents": [{"start": [float(v) for v in ins.guess[a, :3, 0]], "name": f"agent{a}",
                       "goal": [float(v) for v in ins.guess[a, :3, -1]]} for a in range(ins.n_agents)],
           "map": {"dimensions": [int(ins.dimx), int(ins.dimy)],
                   "obstacles": [[float(v) for v in o] for o in ins.obstacles]}}
    with open(path, "w") as f:
        yaml.safe_dump(doc, f)


def test_mapset_driver_round_trip(tmp_path, oracle, params, solver):
    sdir, gdir, odir = tmp_path / "scen", tmp_path / "guess", tmp_path / "out"
    sdir.mkdir(); gdir.mkdir()
    names = []
    for k, seed in enumerate((301, 302, 303)):
        ins = synthetic_instance(seed, 50.0, 4 + k, 10, (8, 14), params, f"map_50by50_obst10_agents{4 + k}_ex{k}")
        _write_scenario(str(sdir / (ins.name + ".yaml")), ins)
        dump_solutions(str(gdir / (ins.name + "_guesses.yaml")), ins.guess, SolutionStatistics())
        names.append(ins.name)
    inst = collect_mapset([str(sdir)], str(gdir))
    assert sorted(i.name for i in inst) == sorted(names)
    rep = run_mapset(inst, solver, str(odir))
    assert len(rep.files) == 3 and rep.refine_seconds > 0
    # oracle on the same instances (planes by the oracle, guesses as read from the files)
    for ins in inst:
        ins.plane_t, ins.plane_abc, _ = oracle.instance_planes(params, ins.guess)
    b = pack_instances(inst)
    ro, _ = oracle.refine(params, b, linsys=0, nthreads=4)
    assert rep.solver_status.tolist() == ro.inst_status.tolist()
    assert rep.success.tolist() == (np.abs(ro.inst_status) <= 2).tolist()
    for i, f in enumerate(rep.files):
        st, ok = read_solution_status(f)
        assert st.solver_status == int(ro.inst_status[i]) and ok == bool(rep.success[i])
        got = load_solutions(f)
        a0, a1 = int(b.inst_agent_ptr[i]), int(b.inst_agent_ptr[i + 1])
        want = np.stack([ro.agent_traj(b, a) for a in range(a0, a1)])
        assert np.abs(got[:, :3] - want[:, :3]).max() <= 1e-3 + 1e-9   # 3-decimal files
    s = rep.summary()
    assert s["instances"] == 3 and 0.0 <= s["success_rate"] <= 1.0
