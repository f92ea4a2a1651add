"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through
the C ABI (ctypes -> libcsdo_dsqp.so), against the CPU oracle on the same
seeded inputs and against the committed golden fixtures.

Bars: bit-exact for the integer/compare work (corridor boxes, plane lists and,
given identical disc centres, plane coefficients; statuses and iteration
counters); trajectories within north_star's 1e-3 m / 1e-3 rad (asserted at
1e-6, observed ~1e-9); the QP objective within OSQP's eps_abs + eps_rel*|obj|.
"""
import os

import numpy as np
import pytest

from conftest import make_batch
from csdotrajectoryplanning_b200 import default_params, pack_instances
from csdotrajectoryplanning_b200.batch import Batch, Instance
from csdotrajectoryplanning_b200.scenario import synthetic_instance

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TRAJ_TOL = 1e-6   # north_star allows 1e-3; the FP64 iterate-following path is far inside it


def _assert_same_refine(ro, rg, traj_tol=TRAJ_TOL):
    for k in ("status", "sqp_iters", "n_qp", "admm_iters", "n_factor", "inst_status", "inst_static_legal"):
        assert np.array_equal(getattr(ro, k), getattr(rg, k)), k
    assert np.abs(ro.traj - rg.traj).max() < traj_tol
    assert np.abs(ro.corridors - rg.corridors).max() < traj_tol
    eps = 1e-3 + 1e-3 * np.abs(ro.objective)
    assert np.all(np.abs(ro.objective - rg.objective) < eps)


def test_planes_bit_exact(oracle, params, solver):
    inst = [synthetic_instance(s, 50.0, 9, 10, (8, 16), params) for s in (31, 32)]
    b0 = pack_instances(inst)
    pb, legal = solver.planes(b0)
    for i, ins in enumerate(inst):
        pts, pabc, lg = oracle.instance_planes(params, ins.guess)
        assert bool(legal[i]) == lg
        for j in range(ins.n_agents):
            a = int(b0.inst_agent_ptr[i]) + j
            k0, k1 = int(pb.plane_ptr[a]), int(pb.plane_ptr[a + 1])
            assert np.array_equal(pb.plane_t[k0:k1], pts[j])
            assert np.array_equal(pb.plane_abc[12 * k0:12 * k1].reshape(-1, 12), pabc[j])
    assert pb.plane_ptr[-1] > 0


def test_planes_detect_initial_collision(oracle, params, solver):
    g = np.zeros((2, 6, 5))
    g[0, 0], g[0, 1] = 20.0, 20.0
    g[1, 0], g[1, 1] = 21.0, 20.5       # overlapping rectangles
    b0 = pack_instances([Instance(g, 50.0, 50.0, np.zeros((0, 3)))])
    pb, legal = solver.planes(b0)
    _, _, lg = oracle.instance_planes(params, g)
    assert not lg and legal[0] == 0 and pb.plane_ptr[-1] == 10


@pytest.mark.parametrize("double_centres", [False, True])
def test_corridors_bit_exact(oracle, params, solver, double_centres):
    b = make_batch(oracle, params, [41, 42], na=6, no=25)
    corr, bs, legal = solver.corridors(b, double_centres)
    for a in range(b.n_agents):
        i = int(np.searchsorted(b.inst_agent_ptr, a, side="right") - 1)
        g, o = b.agent_guess(a), int(b.agent_off[a])
        nt = g.shape[1]
        ob = b.obs[3 * b.obs_ptr[i]:3 * b.obs_ptr[i + 1]].reshape(-1, 3)
        c0, s0, _ = oracle.agent_corridors(params, g[0], g[1], g[2], 50.0, 50.0, ob, double_centres)
        c1 = corr[8 * o:8 * (o + nt)].reshape(8, nt)
        if double_centres:
            # full-double disc centres: CUDA and glibc sin/cos differ in the last ulp,
            # the box is centre +- k*0.1 so the same ulp shows up; the expansion counts must agree
            assert np.abs(c1 - c0).max() < 1e-12
        else:
            assert np.array_equal(c1, c0)   # float-rounded centres: bit-exact
        assert np.array_equal(bs[4 * o:4 * (o + nt)].reshape(nt, 2, 2), s0)


def test_corridors_illegal_starts(oracle, params, solver):
    """Disc centres outside the map / inside an obstacle square: projection and legal-point search."""
    obs = np.array([[25.0, 25.0, 0.8], [30.0, 10.0, 0.8]])
    g = np.zeros((3, 6, 4))
    g[0, 0], g[0, 1] = 0.2, 10.0                     # out of map
    g[1, 0], g[1, 1] = 24.0, 25.3                    # front disc inside the obstacle square
    g[2, 0], g[2, 1], g[2, 2] = 40.0, 40.0, 0.7      # legal
    b = pack_instances([Instance(g, 50.0, 50.0, obs, [np.zeros(0, np.int32)] * 3, [np.zeros((0, 12))] * 3)])
    corr, bs, legal = solver.corridors(b, False)
    assert legal[0] == 0
    for a in range(3):
        c0, s0, _ = oracle.agent_corridors(params, g[a, 0], g[a, 1], g[a, 2], 50.0, 50.0, obs, False)
        assert np.array_equal(corr[8 * 4 * a:8 * 4 * (a + 1)].reshape(8, 4), c0)
        assert np.array_equal(bs[4 * 4 * a:4 * 4 * (a + 1)].reshape(4, 2, 2), s0)
    st = bs.reshape(3, 4, 2, 2)   # [agent][t][front/rear][success, initial_status]
    assert st[0, 0, 1, 1] == 1 and st[1, 0, 0, 1] == 2 and np.all(st[2, :, :, 1] == 0)


@pytest.mark.parametrize("seeds,na,no", [([51, 52, 53], 5, 12), ([61], 10, 25), ([71, 72], 4, 0)])
def test_refine_matches_oracle(oracle, params, solver, seeds, na, no):
    b = make_batch(oracle, params, seeds, na=na, no=no)
    ro, _ = oracle.refine(params, b, linsys=0, nthreads=4)
    rg = solver.refine(b)
    _assert_same_refine(ro, rg)


def test_refine_golden_fixture(params, solver):
    gd = np.load(os.path.join(GOLD, "dsqp_refine_golden.npz"))
    b = Batch(*(gd[k] for k in ("inst_agent_ptr", "inst_nt", "inst_dims", "obs_ptr", "obs", "agent_off",
                                "guess", "plane_ptr", "plane_t", "plane_abc")))
    rg = solver.refine(b)
    for k in ("status", "sqp_iters", "admm_iters", "n_factor", "inst_status", "inst_static_legal"):
        assert np.array_equal(getattr(rg, k), gd[k]), k
    assert np.abs(rg.traj - gd["traj"]).max() < TRAJ_TOL
    assert np.abs(rg.corridors - gd["corridors"]).max() < TRAJ_TOL


@pytest.mark.parametrize("k_iter", [1, 25, 50, 75])
def test_admm_iterates_follow_oracle(oracle, params, k_iter):
    """Same ADMM iterate sequence: stop both after exactly k iterations of the first QP."""
    from csdotrajectoryplanning_b200.solver import DsqpSolver
    p = params.copy()
    p.max_iter, p.osqp_max_iter = 1, k_iter
    b = make_batch(oracle, p, [81], na=6, no=12)
    ro, _ = oracle.refine(p, b, linsys=0, nthreads=2)
    s = DsqpSolver(p)
    try:
        rg = s.refine(b)
    finally:
        s.close()
    assert np.array_equal(ro.status, rg.status) and np.array_equal(ro.admm_iters, rg.admm_iters)
    assert np.array_equal(ro.n_factor, rg.n_factor)
    # iterate-level agreement: the oracle solves the (n+m) KKT system with a sparse LDL', the kernel the reduced
    # band system with partitions + block cyclic reduction (explicit 6x6 inverses): ~1e-9 relative per solve
    assert np.abs(ro.traj - rg.traj).max() < 1e-7


def test_ragged_batch_and_edge_sizes(oracle, params, solver):
    """Instances with different horizons/agent counts in one launch; Nt = 3; one agent; no planes."""
    inst = [synthetic_instance(91, 50.0, 1, 0, (3, 3), params), synthetic_instance(92, 50.0, 7, 25, (12, 20), params),
            synthetic_instance(93, 50.0, 3, 5, (5, 6), params)]
    g = np.zeros((1, 6, 3))
    g[0, 0] = [10.0, 10.5, 11.0]; g[0, 1] = 10.0; g[0, 4, :2] = 0.5 / params.dt
    inst.append(Instance(g, 50.0, 50.0, np.zeros((0, 3))))
    for ins in inst:
        ins.plane_t, ins.plane_abc, _ = oracle.instance_planes(params, ins.guess)
    b = pack_instances(inst)
    ro, _ = oracle.refine(params, b, linsys=0, nthreads=4)
    rg = solver.refine(b)
    _assert_same_refine(ro, rg)


@pytest.mark.parametrize("acts,na,size", [((27, 30), 6, 50.0), ((40, 42), 6, 60.0), ((46, 47), 5, 70.0), ((50, 52), 6, 60.0),
                                          ((80, 80), 3, 100.0), ((135, 140), 2, 100.0)])
def test_long_horizons_cover_every_kernel_variant(oracle, params, solver, acts, na, size):
    """Horizons 82..421: 16 partitions with the two-level separator solve, the 96/128/160/256/512-thread
    kernel variants, row arrays and (for the longest) the band factor in global scratch."""
    inst = [synthetic_instance(400 + acts[0] + k, size, na, 12, acts, params) for k in range(2)]
    for ins in inst:
        ins.plane_t, ins.plane_abc, _ = oracle.instance_planes(params, ins.guess)
    b = pack_instances(inst)
    ro, _ = oracle.refine(params, b, linsys=0, nthreads=8)
    rg = solver.refine(b)
    # iteration / factorization counts identical; trajectories to 1e-5: an agent with 6 SQP iterations and
    # 750 ADMM iterations at Nt = 241 moves by 1.5e-6 between the oracle's own two linear-system paths
    _assert_same_refine(ro, rg, traj_tol=1e-5)
    info = solver.last_launch()
    assert info["block"] >= b.inst_nt.max() and info["ctas_per_sm"] >= 1


def test_statically_illegal_and_infeasible_agents(oracle, params, solver):
    """An agent starting inside an obstacle square (status bookkeeping, solution0 fallback paths)."""
    ins = synthetic_instance(95, 50.0, 4, 6, (8, 12), params)
    g = ins.guess
    obs = np.vstack([ins.obstacles, [[g[0, 0, 3] + 0.4, g[0, 1, 3] + 0.3, 0.8]]])
    ins = Instance(g, 50.0, 50.0, obs)
    ins.plane_t, ins.plane_abc, _ = oracle.instance_planes(params, g)
    b = pack_instances([ins])
    ro, _ = oracle.refine(params, b, linsys=0, nthreads=2)
    rg = solver.refine(b)
    assert ro.inst_static_legal[0] == 0
    _assert_same_refine(ro, rg, traj_tol=1e-5)


def test_device_resident_path_and_determinism(oracle, params, solver):
    """csdo_refine_device on torch tensors == csdo_refine on host buffers, bit for bit, run to run."""
    import torch
    from csdotrajectoryplanning_b200.solver import DeviceBatch, DeviceResult
    b = make_batch(oracle, params, [101, 102], na=5, no=12)
    r1 = solver.refine(b)
    r2 = solver.refine(b)
    db, dr = DeviceBatch(b, "cuda:0"), DeviceResult(b, "cuda:0")
    solver.refine_device(db, dr, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    r3 = dr.to_host()
    for k in ("traj", "corridors", "status", "sqp_iters", "admm_iters", "n_factor", "objective", "inst_status"):
        assert np.array_equal(getattr(r1, k), getattr(r2, k)), k
        assert np.array_equal(getattr(r1, k), getattr(r3, k)), k


def test_instance_order_invariance(oracle, params, solver):
    """Size-independent property: results do not depend on where an instance sits in the batch."""
    b = make_batch(oracle, params, [111, 112, 113], na=4, no=8)
    insts = b.unpack()
    r = solver.refine(b)
    rb = solver.refine(pack_instances(insts[::-1]))
    off = 0
    for i in reversed(range(3)):
        a0, a1 = int(b.inst_agent_ptr[i]), int(b.inst_agent_ptr[i + 1])
        n = a1 - a0
        assert np.array_equal(rb.status[off:off + n], r.status[a0:a1])
        s0, s1 = int(b.agent_off[a0]), int(b.agent_off[a1])
        so = int(pack_instances(insts[::-1]).agent_off[off])
        assert np.array_equal(rb.traj[6 * so:6 * (so + s1 - s0)], r.traj[6 * s0:6 * s1])
        off += n


def test_solver_dsqp_class_mirror(oracle, params):
    """The reference-shaped class: constructor refines, getters and public members as in dsqp_solver.h."""
    from csdotrajectoryplanning_b200.solver import SolverDSQP, find_neighbor_pairs_and_planes
    ins = synthetic_instance(121, 50.0, 5, 10, (8, 12), params)
    planes, legal = find_neighbor_pairs_and_planes(ins.guess, params)
    solutions = []
    s = SolverDSQP(solutions, ins.guess, planes, ins.dimx, ins.dimy, [tuple(o) for o in ins.obstacles], params, 0)
    ins.plane_t, ins.plane_abc = [p[0] for p in planes], [p[1] for p in planes]
    ro, _ = oracle.refine(params, pack_instances([ins]), linsys=0, nthreads=2)
    assert s.getSolverStatus() == ro.inst_status[0] and s.get_initial_static_legal() == bool(ro.inst_static_legal[0])
    assert s.num_iterations == list(ro.sqp_iters) and len(solutions) == 5 and len(s.corridors[0]) == ins.nt
    x = np.array([[r.x for r in row] for row in solutions])
    assert np.abs(x.reshape(-1) - ro.traj.reshape(5, 6, -1)[:, 0].reshape(-1)).max() < TRAJ_TOL
    assert s.getMaxOfRuntimes() > 0


def test_errors_are_reported(params, solver):
    from csdotrajectoryplanning_b200 import binding
    g = np.zeros((1, 6, 2))
    b = pack_instances([Instance(g, 50.0, 50.0, np.zeros((0, 3)), [np.zeros(0, np.int32)], [np.zeros((0, 12))])])
    with pytest.raises(binding.CsdoError) as e:
        solver.refine(b)
    assert e.value.code == binding.CSDO_ERR_INVALID
    g = np.zeros((1, 6, 600)); g[0, 0] = 10; g[0, 1] = 10
    b = pack_instances([Instance(g, 50.0, 50.0, np.zeros((0, 3)), [np.zeros(0, np.int32)], [np.zeros((0, 12))])])
    with pytest.raises(binding.CsdoError) as e:
        solver.refine(b)
    assert e.value.code == binding.CSDO_ERR_UNSUPPORTED


def test_cpp_solver_dsqp_shim(oracle, params, solver, tmp_path):
    """include/csdo/dsqp_solver.h (the reference-shaped C++ class) gives the same result as the C ABI."""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp", "test_solver_dsqp")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.dirname(exe)])
    ins = synthetic_instance(131, 50.0, 4, 8, (8, 12), params)
    na, nt = ins.n_agents, ins.nt
    f = tmp_path / "guess.txt"
    with open(f, "w") as fh:
        fh.write(f"{na} {nt} {ins.dimx!r} {ins.dimy!r} {ins.obstacles.shape[0]}\n")
        for o in ins.obstacles:
            fh.write(" ".join(repr(float(v)) for v in o) + "\n")
        for a in range(na):
            for t in range(nt):
                fh.write(" ".join(repr(float(ins.guess[a, k, t])) for k in range(6)) + "\n")
    out = subprocess.run([exe, str(f)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    # the caller passed a std::unordered_set<Location> (the reference's argument type): same iteration order here
    order = [int(v) for v in lines.pop(0).split()[1:]]
    assert sorted(order) == list(range(ins.obstacles.shape[0]))
    ins.obstacles = ins.obstacles[order]
    pb, legal = solver.planes(pack_instances([ins]))
    rg = solver.refine(pb)
    head = lines[0].split()
    assert int(head[1]) == rg.inst_status[0] and int(head[3]) == rg.inst_static_legal[0] and int(head[5]) == legal[0]
    for a in range(na):
        tok = lines[1 + a].split()
        assert int(tok[2]) == rg.status[a] and int(tok[3]) == rg.sqp_iters[a]
        assert int(tok[4]) == pb.plane_ptr[a + 1] - pb.plane_ptr[a]
        vals = np.array([float(v) for v in tok[5:]]).reshape(nt, 2)
        assert np.array_equal(vals[:, 0], rg.agent_traj(pb, a)[0])          # x(t), bit for bit
        assert np.array_equal(vals[:, 1], rg.agent_corridor(pb, a)[0])      # xf_min(t)
