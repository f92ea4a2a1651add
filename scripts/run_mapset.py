"""Refine every scenario of a map set as one GPU batch.

usage: python scripts/run_mapset.py --scenarios DIR_OR_FILES... --guesses DIR --out DIR [--device 0]
  --scenarios  benchmark scenario YAMLs (reference format: agents/start/goal, map/dimensions/obstacles)
  --guesses    directory with `<scenario>_guesses.yaml` (x0_bar as dumped by the reference, csdo.cc:139)
  --out        one `<scenario>.yaml` per instance in dumpSolutions format + summary.json
"""
import argparse, json, os, sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from csdotrajectoryplanning_b200 import default_params
from csdotrajectoryplanning_b200.driver import collect_mapset, run_mapset
from csdotrajectoryplanning_b200.solver import DsqpSolver

ap = argparse.ArgumentParser()
ap.add_argument("--scenarios", nargs="+", required=True)
ap.add_argument("--guesses", required=True)
ap.add_argument("--out", required=True)
ap.add_argument("--device", type=int, default=0)
ap.add_argument("--dump-corridor", action="store_true", help="also write <name>_corridors.yaml (csdo.cc:163-165)")
a = ap.parse_args()
inst = collect_mapset(a.scenarios, a.guesses)
if not inst:
    sys.exit("no scenario with a matching guess file")
rep = run_mapset(inst, DsqpSolver(default_params(), device=a.device), a.out, dump_corridor=a.dump_corridor)
s = rep.summary()
json.dump({**s, "per_instance": [{"name": n, "solver_status": int(st), "search_status": int(ss), "collisions": int(c)}
                                 for n, st, ss, c in zip(rep.names, rep.solver_status, rep.search_status, rep.collisions)]},
          open(os.path.join(a.out, "summary.json"), "w"), indent=1)
print(json.dumps(s))
