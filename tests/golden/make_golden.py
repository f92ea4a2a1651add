"""Mint the golden fixtures of tests/golden/ (test infrastructure).

The reference tree ships no golden vector for the DSQP path and neither the
reference nor OSQP 0.6.3 can be built offline, so these fixtures are produced
by the CPU oracle (oracle/, linsys 0 = KKT LDL^T) on seeded synthetic
instances.  Regenerate with:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from csdotrajectoryplanning_b200 import default_params, pack_instances  # noqa: E402
from csdotrajectoryplanning_b200.scenario import synthetic_instance  # noqa: E402
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    p = default_params()
    inst = [synthetic_instance(2024, 50.0, 6, 25, (10, 18), p, "golden_obst"),
            synthetic_instance(2025, 50.0, 4, 0, (8, 14), p, "golden_empty")]
    for ins in inst:
        ins.plane_t, ins.plane_abc, _ = O.instance_planes(p, ins.guess)
    b = pack_instances(inst)
    res, _ = O.refine(p, b, linsys=0, nthreads=1)
    np.savez_compressed(
        os.path.join(HERE, "dsqp_refine_golden.npz"),
        inst_agent_ptr=b.inst_agent_ptr, inst_nt=b.inst_nt, inst_dims=b.inst_dims, obs_ptr=b.obs_ptr,
        obs=b.obs, agent_off=b.agent_off, guess=b.guess, plane_ptr=b.plane_ptr, plane_t=b.plane_t,
        plane_abc=b.plane_abc, traj=res.traj, corridors=res.corridors, status=res.status,
        sqp_iters=res.sqp_iters, admm_iters=res.admm_iters, n_factor=res.n_factor,
        objective=res.objective, inst_status=res.inst_status, inst_static_legal=res.inst_static_legal)
    # corridor boxes: hand-checkable cases + the accumulated-0.1 edge
    obs = np.array([[10.0, 10.0, 0.8], [20.0, 12.0, 0.8], [14.0, 22.0, 0.5]])
    pts = np.array([[25.0, 25.0], [12.5, 10.2], [1.0, 30.0], [10.3, 10.1], [49.5, 49.9], [16.0, 14.0]])
    boxes, stats = [], []
    for (x, y) in pts:
        bx, st = O.generate_box(p, 50.0, 50.0, float(x), float(y), obs)
        boxes.append(bx); stats.append(st)
    np.savez_compressed(os.path.join(HERE, "corridor_boxes_golden.npz"), obs=obs, pts=pts,
                        boxes=np.asarray(boxes), status=np.asarray(stats))
    print("golden fixtures written:", res.status, res.sqp_iters, res.admm_iters)


if __name__ == "__main__":
    main()
