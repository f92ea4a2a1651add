"""CPU test of the map-set driver's host logic: file pairing, batch assembly,
solution files, summary.  The refine itself is stood in by the CPU oracle here (tests may use it); the
GPU counterpart is tests/test_gpu_driver.py."""
import json
import os

import numpy as np
import yaml

from csdotrajectoryplanning_b200 import pack_instances
from csdotrajectoryplanning_b200.driver import collect_mapset, run_mapset
from csdotrajectoryplanning_b200.output import SolutionStatistics, dump_solutions, load_solutions, read_solution_status
from csdotrajectoryplanning_b200.scenario import synthetic_instance


class _OracleSolver:
    """Same two calls the driver makes on DsqpSolver."""

    def __init__(self, oracle, params):
        self.o, self.p = oracle, params
        self.params = params

    def planes(self, batch):
        inst = batch.unpack()
        legal = np.ones(batch.n_inst, np.int32)
        for i, ins in enumerate(inst):
            ins.plane_t, ins.plane_abc, ok = self.o.instance_planes(self.p, ins.guess)
            legal[i] = int(ok)
        return pack_instances(inst), legal

    def refine(self, batch):
        return self.o.refine(self.p, batch, linsys=1, nthreads=2)[0]


def test_mapset_driver_host_logic(tmp_path, oracle, params):
    sdir, gdir, odir = tmp_path / "scen", tmp_path / "guess", tmp_path / "out"
    sdir.mkdir(); gdir.mkdir()
    for k, seed in enumerate((311, 312)):
        ins = synthetic_instance(seed, 50.0, 3 + k, 6, (6, 9), params, f"map_50by50_obst6_agents{3 + k}_ex{k}")
        doc = {"agents": [{"start": [float(v) for v in ins.guess[a, :3, 0]], "name": f"agent{a}",
                           "goal": [float(v) for v in ins.guess[a, :3, -1]]} for a in range(ins.n_agents)],
               "map": {"dimensions": [50, 50], "obstacles": [[float(v) for v in o] for o in ins.obstacles]}}
        with open(sdir / (ins.name + ".yaml"), "w") as f:
            yaml.safe_dump(doc, f)
        dump_solutions(str(gdir / (ins.name + "_guesses.yaml")), ins.guess, SolutionStatistics())
    # a scenario without a guess file is skipped
    with open(sdir / "orphan.yaml", "w") as f:
        yaml.safe_dump({"agents": [], "map": {"dimensions": [50, 50], "obstacles": []}}, f)
    inst = collect_mapset([str(sdir)], str(gdir))
    assert [i.name for i in inst] == ["map_50by50_obst6_agents3_ex0", "map_50by50_obst6_agents4_ex1"]
    assert inst[0].guess.shape[0] == 3 and inst[1].guess.shape[0] == 4 and inst[0].obstacles.shape == (6, 3)
    rep = run_mapset(inst, _OracleSolver(oracle, params), str(odir), dump_corridor=True)
    assert len(rep.files) == 2 and all(os.path.exists(f) for f in rep.files)
    # --dump_corridor files: "agent<a>:" + two rows per step, readable as YAML
    cdoc = yaml.safe_load(open(rep.files[0][:-5] + "_corridors.yaml"))
    assert sorted(cdoc) == ["agent0", "agent1", "agent2"] and len(cdoc["agent0"]) == 2 * inst[0].guess.shape[2]
    row = cdoc["agent1"][0]
    assert len(row) == 6 and row[2] <= row[0] <= row[3] and row[4] <= row[1] <= row[5]   # the disc centre is inside its box
    for i, f in enumerate(rep.files):
        st, ok = read_solution_status(f)
        assert st.solver_status == int(rep.solver_status[i]) and ok == bool(rep.success[i])
        assert st.search_status == int(rep.search_status[i])
        assert load_solutions(f).shape == inst[i].guess.shape
    s = rep.summary()
    assert s["instances"] == 2 and 0.0 <= s["success_rate"] <= 1.0 and len(rep.collisions) == 2
    json.dumps(s)
