"""CPU tests of the host logic: batch packing, parameters, the C-ABI library's
exports, sharding (incl. a world_size-2 gloo run).  No compute call needs a GPU."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from csdotrajectoryplanning_b200 import default_params, pack_instances
from csdotrajectoryplanning_b200 import sharding
from csdotrajectoryplanning_b200.batch import RefineResult
from csdotrajectoryplanning_b200.params import CsdoParams

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_default_params_match_reference_config():
    p = default_params()
    assert p.f2x == 1.25 and p.r2x == -0.25 and p.rv == 1.25 and p.WB == 1.0
    # dt = (double)(float(3)*float(0.706)) / 1 / 3 / 0.8 (sqp/utils.cc:55-56)
    assert p.dt == 0.8825000127156575
    assert p.steer_max == np.arctan(1.0 / 3.0)
    assert (p.max_iter, p.osqp_max_iter, p.adaptive_rho_interval, p.scaling) == (10, 400, 25, 10)


def test_library_exports_every_declared_symbol():
    """include/*.h <-> libcsdo_dsqp.so: every declared entry point is exported."""
    from csdotrajectoryplanning_b200 import binding
    hdr = open(os.path.join(ROOT, "include", "csdo_dsqp.h")).read()
    declared = sorted(set(re.findall(r"\b(csdo_[a-z0-9_]+)\s*\(", hdr)))
    assert set(declared) == set(binding.EXPORTS)
    L = binding.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.csdo_version().startswith(b"csdo-dsqp-b200")
    # parameters computed in C == parameters computed on the host side
    cp = CsdoParams()
    L.csdo_default_params(C.byref(cp))
    hp = default_params()
    for name, _ in CsdoParams._fields_:
        assert getattr(cp, name) == getattr(hp, name), name


def test_no_gpu_means_loud_failure():
    """The product path never falls back to the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from csdotrajectoryplanning_b200 import binding
    from csdotrajectoryplanning_b200.solver import DsqpSolver
    with pytest.raises(binding.CsdoError) as e:
        DsqpSolver(default_params())
    assert e.value.code == binding.CSDO_ERR_CUDA


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "csdotrajectoryplanning_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dp, f)).read()
                bad = re.findall(r"^\s*(?:from|import)\s+oracle\b|#include\s*[\"<][^\">]*oracle|dlopen|libdsqp_oracle",
                                 src, flags=re.M)
                assert not bad, f"{f} reaches into the oracle: {bad}"
    out = subprocess.run(["ldd", os.path.join(pkg, "csrc", "libcsdo_dsqp.so")], capture_output=True, text=True)
    assert "dsqp_oracle" not in out.stdout


def test_batch_roundtrip(small_batch):
    b = small_batch
    b.validate()
    b2 = pack_instances(b.unpack())
    for k in ("inst_agent_ptr", "inst_nt", "inst_dims", "obs_ptr", "obs", "agent_off", "guess",
              "plane_ptr", "plane_t", "plane_abc"):
        assert np.array_equal(getattr(b, k), getattr(b2, k)), k
    sub = b.select_instances([1])
    assert sub.n_inst == 1 and sub.n_agents == b.inst_agent_ptr[2] - b.inst_agent_ptr[1]


def test_split_balanced():
    cost = np.array([5, 1, 1, 1, 5, 5, 1, 1], np.int64)
    for parts in (1, 2, 3, 4, 8, 16):
        r = sharding.split_balanced(cost, parts)
        assert len(r) == parts and r[0][0] == 0 and r[-1][1] == len(cost)
        assert all(r[i][1] == r[i + 1][0] for i in range(parts - 1))
    assert sharding.split_balanced(cost, 2) == [(0, 4), (4, 8)]


def test_instance_and_agent_sharding_cover_everything(small_batch):
    b = small_batch
    for world in (1, 2, 3):
        seen_i, seen_a = [], []
        for r in range(world):
            sb, (i0, i1) = sharding.shard_instances(b, r, world)
            seen_i += list(range(i0, i1))
            assert sb.n_inst == i1 - i0
            pb, ids = sharding.shard_agents(b, r, world)
            assert pb.n_inst == b.n_inst and pb.n_agents == len(ids)
            seen_a += list(ids)
            for j, a in enumerate(ids):
                assert np.array_equal(pb.agent_guess(j), b.agent_guess(int(a)))
                k0, k1 = b.plane_ptr[a], b.plane_ptr[a + 1]
                assert np.array_equal(pb.plane_t[pb.plane_ptr[j]:pb.plane_ptr[j + 1]], b.plane_t[k0:k1])
        assert seen_i == list(range(b.n_inst)) and sorted(seen_a) == list(range(b.n_agents))


def test_agent_partition_equals_whole_refine(oracle, params, small_batch):
    """Agents are independent during DSQP (planes are frozen): partitioned == whole, bit for bit."""
    b = small_batch
    whole, _ = oracle.refine(params, b, linsys=1, nthreads=2)
    fn = lambda pb: oracle.refine(params, pb, linsys=1, nthreads=1)[0]
    full = RefineResult.allocate(b)
    legal = np.ones(b.n_inst, np.int32)
    for r in range(2):
        pb, ids = sharding.shard_agents(b, r, 2)
        part = fn(pb)
        sharding.scatter_agent_results(full, b, ids, part, pb)
        legal &= part.inst_static_legal
    full.inst_static_legal[:] = legal
    sharding.aggregate_instance_status(b, full)
    for k in ("traj", "corridors", "status", "sqp_iters", "admm_iters", "n_factor", "inst_status",
              "inst_static_legal"):
        assert np.array_equal(getattr(full, k), getattr(whole, k)), k


_GLOO_WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, {root!r})
sys.path.insert(0, os.path.join({root!r}, "tests"))
from conftest import make_batch
from csdotrajectoryplanning_b200 import default_params, sharding
from oracle import oracle as O
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
p = default_params()
b = make_batch(O, p, [11, 12, 13])
fn = lambda pb: O.refine(p, pb, linsys=1, nthreads=1)[0]   # the oracle stands in for the CUDA path on CPU
full = sharding.refine_agent_partitioned(b, fn, rank, world, dist)
whole = fn(b)
ok = all(np.array_equal(getattr(full, k), getattr(whole, k)) for k in
         ("traj", "corridors", "status", "sqp_iters", "admm_iters", "n_factor", "inst_status", "inst_static_legal"))
# instance sharding: no collective, ranks own disjoint instance ranges
sb, (i0, i1) = sharding.shard_instances(b, rank, world)
part = fn(sb)
a0 = int(b.inst_agent_ptr[i0]); a1 = int(b.inst_agent_ptr[i1])
ok = ok and np.array_equal(part.status, whole.status[a0:a1]) and np.array_equal(
    part.traj, whole.traj[6 * int(b.agent_off[a0]):6 * int(b.agent_off[a1])])
dist.barrier()
print("RANK", rank, "OK" if ok else "MISMATCH", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
"""


def test_gloo_world2_agent_partition(tmp_path, oracle):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2",
               OMP_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"RANK {r} OK" in o, o[-2000:]


_GLOO_GATHER_WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, {root!r})
sys.path.insert(0, os.path.join({root!r}, "tests"))
from conftest import make_batch
from csdotrajectoryplanning_b200 import default_params, sharding
from csdotrajectoryplanning_b200.batch import RefineResult
from oracle import oracle as O
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
p = default_params()
b = make_batch(O, p, [21, 22, 23], na=7)          # 7 agents per instance: uneven split over 2 ranks
whole = O.refine(p, b, linsys=1, nthreads=1)[0]   # the oracle stands in for the CUDA path on CPU
ids = sharding.rank_agent_ids(b, rank, world)
# what a rank holds after refining ONLY its own agents: everything else untouched (zeros / stale)
t = {{k: torch.zeros_like(torch.from_numpy(np.ascontiguousarray(getattr(whole, k)))) for k in
     ("traj", "corridors", "status", "sqp_iters", "n_qp", "admm_iters", "n_factor", "objective", "inst_status", "inst_static_legal")}}
for a in ids:
    o0, o1 = int(b.agent_off[a]), int(b.agent_off[a + 1])
    t["traj"][6 * o0:6 * o1] = torch.from_numpy(whole.traj[6 * o0:6 * o1])
    t["corridors"][8 * o0:8 * o1] = torch.from_numpy(whole.corridors[8 * o0:8 * o1])
    for k in ("status", "sqp_iters", "n_qp", "admm_iters", "n_factor", "objective"):
        t[k][a] = float(getattr(whole, k)[a]) if k == "objective" else int(getattr(whole, k)[a])
t["inst_static_legal"][:] = 1
if rank == 1: t["inst_static_legal"][0] = 0       # one rank saw an illegal start in instance 0
g = sharding.DeviceAllGather(b, rank, world, torch.device("cpu"), dist)
g.run(t)
full = RefineResult.allocate(b)
for k in ("traj", "corridors", "status", "sqp_iters", "n_qp", "admm_iters", "n_factor", "objective"):
    getattr(full, k)[:] = t[k].numpy()
sharding.aggregate_instance_status(b, full)
ok = all(np.array_equal(getattr(full, k), getattr(whole, k)) for k in
         ("traj", "corridors", "status", "sqp_iters", "n_qp", "admm_iters", "n_factor", "objective", "inst_status"))
ok = ok and t["inst_static_legal"].tolist() == [0] + [1] * (b.n_inst - 1)
covered = np.concatenate([sharding.rank_agent_ids(b, r, world) for r in range(world)])
ok = ok and sorted(covered.tolist()) == list(range(b.n_agents))
dist.barrier()
print("RANK", rank, "OK" if ok else "MISMATCH", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
"""


def test_gloo_world2_device_allgather(tmp_path, oracle):
    """sharding.DeviceAllGather (the pack -> all-gather -> unpack of the agent-partitioned GPU mode) on CPU
    tensors over gloo: every rank ends with the whole solution, the legality flag is AND-reduced."""
    script = tmp_path / "worker_gather.py"
    script.write_text(_GLOO_GATHER_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541", WORLD_SIZE="2", OMP_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"RANK {r} OK" in o, o[-2000:]
