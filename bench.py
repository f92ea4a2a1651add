#!/usr/bin/env python
"""Benchmark of the DSQP refine path (BASELINE.json: batched agent-QP solves/sec; refine ms/instance).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5|map50|map100_a100|room]
                    [--instances M] [--partition instances|agents] [--impl reference]

A "step" is one refine pass of the hot path over the whole job (one batch per GPU).  The job is FIXED
(strong scaling): M seeded synthetic instances of the named BASELINE.json configuration; rank r of N
refines instances r, r+N, r+2N, ... (instance sharding, no data-path collective) or, with
`--partition agents`, the agents [a0, a1) of EVERY instance followed by one NCCL all-gather of the
results (BASELINE configs[2]).  The coarse plans are synthetic priority-style plans over the planner's own
motion primitives (the PBS + Hybrid-A* front end is out of scope and cannot be built offline).

`value`  : whole-job QP/s, inputs (guess, obstacles, planes) resident in HBM, CUDA events on the launching
           stream, max over ranks.
`e2e`    : the same metric through csdo_refine() with pinned HOST buffers, H2D + D2H inside the timed region.
`roofline`, `cpu_baseline`, `parity`, `latency`, `preprocess`: see DESIGN.md section 5.
`--impl reference`: the reference's CPU path.  sqp/dsqp_solver.cc needs Eigen + OSQP 0.6.3 and does not build
offline, so this arm times the CPU oracle port (oracle/, OpenMP over agents, all host threads) on a bounded
sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "batched agent-QP solves/sec"
UNIT = "QP/s"

WORKLOADS = {
    # name: (default instances, description) -- tools/synth.py::workload_jobs builds instance i
    "c5": (1024, "BASELINE configs[4] shape: synthetic 100x100 maps, 100 agents (largest benchmark agent count), "
                 "50 obstacles, horizons 127/190/256 in turn"),
    "map50": (600, "BASELINE configs[1] shape: map50by50 sweep, agents 5/10/15/20/25 x {empty, 25 obstacles}"),
    "map100_a100": (60, "BASELINE configs[2] shape: map100by100 / agents100 / obstacle (100 agents, 50 obstacles)"),
    "room": (300, "BASELINE configs[3] shape: room maps 100x100, agents 10..50, 130..298 wall discs of r = 0.5"),
    "real": (579, "REAL benchmark geometry: the 579 scenarios of benchmark/{map50by50,map100by100,room} that the stand-in "
                  "planner (tools/coarse_planner.cpp) routes completely (tests/golden/real_scenarios.npz: real maps, "
                  "obstacles, starts and goals; coarse plans by the stand-in, x0_bar by InterpolateInitalGuess)"),
}
C5_FULL = 4096   # BASELINE configs[4] names 4096 concurrent instances


def workload_string(name: str, total: int) -> str:
    s = f"{WORKLOADS[name][1]}; {total} instances per step (whole job, split over the GPUs)"
    if name == "c5" and total != C5_FULL:
        s += (f"; configs[4] names {C5_FULL} instances -- {total} keeps one step within seconds "
              f"(--instances {C5_FULL} runs the full size)")
    return s + ("" if name == "real" else "; synthetic priority-style plans")


def build_instances(name: str, total: int, rank: int, world: int, params):
    if name == "real":     # instance i = scenario i mod N of the committed fixture
        from csdotrajectoryplanning_b200.driver import instances_from_coarse_plans
        fix = os.path.join(ROOT, "tests", "golden", "real_scenarios.npz")
        n_fix = len(np.load(fix)["name"])
        return instances_from_coarse_plans(fix, params, select=[i % n_fix for i in range(total)][rank::world])
    from tools import synth
    jobs = synth.workload_jobs(name, total)[rank::world]
    return synth.synth_jobs(jobs, params)


def algorithmic_flops(batch, res) -> float:
    """F_QP = n_fac*504*Nt + n_it*(518*Nt + 88*K) + floor(n_it/25)*F_chk + F_scale,
    evaluated with the COUNTED iterations/factorizations of every agent (summed over its QPs).
    res: anything with admm_iters / n_factor / n_qp arrays."""
    nt = batch.agent_nt().astype(np.float64)
    K = np.diff(batch.plane_ptr).astype(np.float64)
    nnzA, nnzP, n, m = 28 * nt - 11 + 12 * K, 5 * nt, 6 * nt - 2, 13 * nt + 4 * K
    f_chk = 6 * nnzA + 2 * nnzP + 8 * (n + m)
    f_scale = 10 * 4 * (nnzA + nnzP)
    g = (lambda k: np.asarray(res[k] if isinstance(res, dict) else getattr(res, k), np.float64))
    it, fac, nqp = g("admm_iters"), g("n_factor"), g("n_qp")
    f = fac * 504 * nt + it * (518 * nt + 88 * K) + np.floor(it / 25) * f_chk + nqp * f_scale
    return float(f.sum())


def algorithmic_bytes(batch) -> float:
    """Compulsory HBM bytes of one refine: inputs once, outputs once (~160 Nt + 104 K per agent)."""
    nt = batch.agent_nt().astype(np.float64)
    K = np.diff(batch.plane_ptr).astype(np.float64)
    return float((8 * (6 * nt + 6 * nt + 8 * nt) + 100 * K + 32).sum())


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[j] for r in self.rows if len(r) >= 8 for j in range(4) if r[4 + j].startswith("Active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU legs (the oracle is loaded ONLY here: cpu_baseline / parity / --impl reference)
def cpu_port_run(p, batch, nthreads: int, linsys: int = 0):
    """The CPU oracle port on `batch` (linsys 0: KKT LDL^T = the cost profile of OSQP/QDLDL, 1: banded)."""
    from oracle import oracle as O
    t0 = time.perf_counter()
    res, _ = O.refine(p, batch, linsys=linsys, nthreads=nthreads)
    return time.perf_counter() - t0, res


def with_oracle_planes(p, inst):
    from oracle import oracle as O
    for ins in inst:
        ins.plane_t, ins.plane_abc, _ = O.instance_planes(p, ins.guess)
    return inst


def cpu_sample(p, inst, target_core_s: float, nthreads: int):
    """Bounded sample: instances spread over the workload worth about target_core_s core-seconds."""
    from csdotrajectoryplanning_b200 import pack_instances
    probe_i = list(range(0, len(inst), max(1, len(inst) // 3)))[:3]
    probe = pack_instances(with_oracle_planes(p, [inst[i] for i in probe_i]))
    t, _ = cpu_port_run(p, probe, nthreads)
    per_agent = t * min(nthreads, probe.n_agents) / max(1, probe.n_agents)      # core-seconds per agent
    want = max(1, int(target_core_s / max(per_agent, 1e-9)))
    chosen, tot = [], 0
    stride = max(1, len(inst) // 64)
    order = list(range(0, len(inst), stride)) + [i for i in range(len(inst)) if i % stride]
    for i in order:
        if tot >= want and chosen:
            break
        chosen.append(i); tot += inst[i].n_agents
    chosen.sort()
    return pack_instances(with_oracle_planes(p, [inst[i] for i in chosen])), chosen


def config_dict(args, total: int) -> dict:
    """The `config` object, identical in both arms (the driver compares them): what a step is."""
    return {"workload": workload_string(args.workload, total), "instances_per_step": total,
            "partition": args.partition,
            "l2": "b200 arm: 256 MiB buffer written between timed iterations (outside the events); "
                  "reference arm: host cores, no GPU"}


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    from csdotrajectoryplanning_b200 import default_params
    p = default_params()
    total = args.instances or WORKLOADS[args.workload][0]
    # the sample is drawn from rank 0's share of the N = 1 job: the same instances whatever N is
    inst = build_instances(args.workload, total, 0, max(1, total // 64), p)
    cores = os.cpu_count() or 1
    budget = 150.0 / max(1, args.steps + args.warmup)          # whole run within a few minutes
    sample, chosen = cpu_sample(p, inst, min(20.0, budget) * cores * 0.7, cores)
    for _ in range(args.warmup):
        cpu_port_run(p, sample, cores)
    t_tot, qps = 0.0, 0
    for _ in range(args.steps):
        t, r = cpu_port_run(p, sample, cores)
        t_tot += t; qps += int(r.n_qp.sum())
    value = qps / t_tot
    desc = (f"{len(chosen)} instances ({sample.n_agents} agents) of the workload per step, all {cores} host threads "
            f"(OpenMP over agents; the reference itself runs the agents sequentially on one core)")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "real benchmark scenarios, stand-in coarse plans" if args.workload == "real" else "synthetic",
            "refine_ms_per_instance": 1e3 * t_tot / args.steps / len(chosen),
            "config": config_dict(args, total), "details": {"sample": desc},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "sqp/dsqp_solver.cc does not build offline (Eigen/OSQP 0.6.3 absent): this is the CPU oracle port of "
                    "its OSQP path (KKT LDL^T like QDLDL); oracle/_ref pins the corridor and plane code only"}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def latency_block(p, solver_cls, device: int) -> dict:
    """BASELINE's second metric for ONE configs[0]-shaped instance (50x50, 25 agents, 25 obstacles): wall
    time of SolverDSQP-equivalent calls (host buffers in, results out), handle kept vs created per call,
    next to the CPU port on one core (what ./csdo does, dsqp_solver.cc:1198) and on all cores."""
    from csdotrajectoryplanning_b200 import pack_instances
    from tools import synth
    ins = synth.synth_jobs([(777, 50.0, 25, 25, (12, 30), "configs0_shape", 0.8)], p)
    S = solver_cls(p, device=device)
    b, _ = S.planes(pack_instances(ins))
    S.refine(b)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); r = S.refine(b); ts.append(time.perf_counter() - t0)
    S.close()
    tf = []
    for _ in range(3):
        t0 = time.perf_counter(); S2 = solver_cls(p, device=device); S2.refine(b); S2.close()
        tf.append(time.perf_counter() - t0)
    out = {"instance": "configs[0] shape: 50x50, 25 agents, 25 obstacles, Nt %d, %d planes" % (b.inst_nt[0], b.plane_ptr[-1]),
           "gpu_ms_handle_reused": 1e3 * float(np.median(ts)), "gpu_ms_fresh_handle": 1e3 * float(np.median(tf)),
           "qps": int(r.n_qp.sum()), "admm_iters": int(r.admm_iters.sum())}
    try:
        cores = os.cpu_count() or 1
        ob = pack_instances(with_oracle_planes(p, ins))
        t1, r1 = cpu_port_run(p, ob, 1)
        tn, _ = cpu_port_run(p, ob, min(cores, 25))
        out.update({"cpu_port_ms_1core": 1e3 * t1, "cpu_port_ms_allcores": 1e3 * tn,
                    "status_equal": bool(np.array_equal(r1.status, r.status)),
                    "max_abs_traj": float(np.abs(r1.traj - r.traj).max())})
    except Exception as e:      # the oracle is a checker: its absence must not fail the bench line
        out["cpu_port_error"] = str(e)[:200]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--instances", type=int, default=0, help="instances per step over all GPUs (0: workload default)")
    ap.add_argument("--partition", default="instances", choices=["instances", "agents"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from csdotrajectoryplanning_b200 import binding, default_params, pack_instances
    from csdotrajectoryplanning_b200.batch import Batch, RefineResult
    from csdotrajectoryplanning_b200.solver import DeviceBatch, DeviceResult, DsqpSolver

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    p = default_params()
    total = args.instances or WORKLOADS[args.workload][0]
    agents_mode = args.partition == "agents"
    t_gen = time.perf_counter()
    if agents_mode:
        from csdotrajectoryplanning_b200 import sharding
        inst = build_instances(args.workload, total, 0, 1, p)      # every rank holds every instance ...
    else:
        inst = build_instances(args.workload, total, rank, world, p)
    t_gen = time.perf_counter() - t_gen
    batch0 = pack_instances(inst)
    n_inst = batch0.n_inst
    solver = DsqpSolver(p, device=local_rank)     # raises without a B200: no CPU fallback

    # a dedicated non-default stream: the kernels and the timing events share it (stream handle 0 would mean
    # "the handle's own stream" to the C ABI)
    stream = torch.cuda.Stream(device=dev)

    # ---- pre-process (a14/a15) on the device: neighbour pairs + planes; timed separately ----
    db = DeviceBatch(batch0, dev, order=False)
    if agents_mode:                                                # ... and owns a slice of each one's agents
        db.set_active(sharding.rank_agent_ids(batch0, rank, world))
    pre_ms = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream); solver.planes_device(db, stream.cuda_stream); e1.record(stream)
        torch.cuda.synchronize(); pre_ms.append(e0.elapsed_time(e1))
    batch = db.planes_to_host()      # (agents mode: plane lists of the other ranks' agents are empty)
    dr = DeviceResult(batch, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    gather = None
    if agents_mode and world > 1:
        gather = sharding.DeviceAllGather(batch0, rank, world, dev, dist)
    evg = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]

    def one_step(e=None, eg=None):
        with torch.cuda.stream(stream):
            flush.fill_(1)                  # L2 flush between iterations (outside the events)
            if e: e[0].record(stream)
            solver.refine_device(db, dr, stream.cuda_stream)
            if gather is not None:
                if eg: eg[0].record(stream)
                gather.run(dr.t)            # pack -> all-gather (f64 + i32) -> unpack, on this stream
                solver.aggregate_status_device(db, dr, stream.cuda_stream)
                if eg: eg[1].record(stream)
            if e: e[1].record(stream)

    for _ in range(args.warmup):
        one_step()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    torch.cuda.synchronize()
    for k in range(args.steps):
        one_step(ev[k], evg[k])
    torch.cuda.synchronize()
    solver.sync()
    if world > 1: dist.barrier()
    clk = clocks.stop()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    gather_ms = [a.elapsed_time(b) for a, b in evg] if gather is not None else []
    t_dev = sum(step_ms) * 1e-3
    cnt = dr.counters_to_host()
    own = db.active_ids.astype(np.int64) if agents_mode else np.arange(batch.n_agents)
    qps_step = int(cnt["n_qp"][own].sum())
    launch = solver.last_launch()

    # ---- end-to-end arm: csdo_refine with pinned host buffers ----
    def pinned_like(a):
        t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype).pin_memory()
        v = t.numpy(); v[...] = a
        return t, v
    keep, hb = [], {}
    for name in ("inst_agent_ptr", "inst_nt", "inst_dims", "obs_ptr", "obs", "agent_off", "guess", "plane_ptr",
                 "plane_t", "plane_abc"):
        t, v = pinned_like(getattr(batch, name)); keep.append(t); hb[name] = v
    hbatch = Batch(**hb)
    if agents_mode:
        hbatch.agent_order, hbatch.n_active = np.ascontiguousarray(db.active_ids, np.int32), int(db.active_ids.shape[0])
    hres = RefineResult.allocate(batch)
    for name in ("traj", "corridors"):
        t, v = pinned_like(getattr(hres, name)); keep.append(t); setattr(hres, name, v)
    h2d = sum(int(v.nbytes) for v in hb.values()) + 4 * batch.n_agents
    d2h = sum(int(getattr(hres, n).nbytes) for n in ("traj", "corridors", "status", "sqp_iters", "n_qp", "admm_iters",
                                                    "n_factor", "objective", "inst_status", "inst_static_legal"))
    solver.refine(hbatch, hres)
    torch.cuda.synchronize()
    # as many of the K steps as fit ~60 s (a c5 step takes seconds); at least 2
    e2e_steps = int(max(2, min(args.steps, 60.0 / max(np.mean(step_ms) * 1e-3, 1e-3))))
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        solver.refine(hbatch, hres)         # blocks until the results are back in host memory
    torch.cuda.synchronize()
    t_e2e = (time.perf_counter() - t0) * args.steps / e2e_steps      # scaled to K steps for the reduction below
    e2e_ok = bool(np.array_equal(hres.status[own], cnt["status"][own]) and
                  np.array_equal(hres.admm_iters[own], cnt["admm_iters"][own]))

    # ---- reduce over ranks: max time, summed work ----
    tot_qp, tot_inst, tot_agents = qps_step * args.steps, (0 if agents_mode and rank else n_inst), int(own.shape[0])
    admm_tot = float(cnt["admm_iters"][own].sum())
    if world > 1:
        tt = torch.tensor([t_dev, t_e2e, max(pre_ms)], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_e2e = float(tt[0]), float(tt[1])
        cc = torch.tensor([tot_qp, tot_inst, tot_agents, admm_tot, h2d, d2h], dtype=torch.float64, device=dev)
        dist.all_reduce(cc, op=dist.ReduceOp.SUM)
        tot_qp, tot_inst, tot_agents, admm_tot = int(cc[0]), int(cc[1]), int(cc[2]), float(cc[3])
        h2d, d2h = int(cc[4]), int(cc[5])
    bit_identical = None
    if gather is not None:      # the gathered result must equal an unsharded refine of the same batch, bit for bit
        db1 = DeviceBatch(batch0, dev, order=False)
        dr1 = DeviceResult(batch0, dev)
        with torch.cuda.stream(stream):
            solver.planes_device(db1, stream.cuda_stream)
            solver.refine_device(db1, dr1, stream.cuda_stream)
        torch.cuda.synchronize()
        same = all(bool(torch.equal(dr.t[k], dr1.t[k])) for k in dr.t)
        flag = torch.tensor([1.0 if same else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        bit_identical = bool(flag.item() == 1.0)
    if rank != 0:
        if world > 1: dist.destroy_process_group()
        return

    value = tot_qp / t_dev
    # ---- roofline of the dominant kernel (dsqp_refine_kernel): FP64 pipe, measured live ----
    if agents_mode:      # rank 0's own agents only (the other agents' counters arrived by all-gather)
        mask = np.zeros(batch.n_agents, bool); mask[own] = True
        cnt = {k: (np.where(mask, v, 0) if v.shape[0] == batch.n_agents else v) for k, v in cnt.items()}
    fl = algorithmic_flops(batch, cnt)
    by = algorithmic_bytes(batch) * (own.shape[0] / max(batch.n_agents, 1))
    kern_s = (float(np.mean(step_ms)) - (float(np.mean(gather_ms)) if gather_ms else 0.0)) * 1e-3
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    import ctypes as C
    fp64 = C.c_double(0.0)
    binding.lib().csdo_measure_fp64_peak(local_rank, C.byref(fp64))
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = f"dram_bytes_per_agent_step_{args.workload}"
        if key in tj:       # ncu capture of a sub-batch of this workload, scaled by agent steps
            traffic = tj[key] * float(batch.agent_nt()[own].sum())
            traffic_src = f'{tj.get("source")} (git {tj.get("git")}): ncu DRAM bytes per agent step x the agent steps of this launch'
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    roofline = {"bound": "fp64", "achieved": fl / kern_s * 1e-12, "peak": fp64.value, "unit": "TFLOP/s",
                "frac": (fl / kern_s * 1e-12) / fp64.value if fp64.value else None, "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": "builder-measured: csdo_measure_fp64_peak (DFMA microbenchmark, this run); FP64 is not in "
                               "MEASURED_PEAKS.json",
                "algorithmic_flops_per_launch": fl, "kernel_ms": 1e3 * kern_s,
                "hbm": {"bound": "hbm", "achieved": by / kern_s * 1e-9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": by / kern_s * 1e-9 / hbm_peak, "algorithmic_bytes_per_launch": by,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)"},
                "launches_per_step": 1,
                "note": "banded FP64 ADMM with the working set in shared memory: neither HBM nor tensor cores bind it; "
                        "rank 0's share of the job; one step = one persistent launch of dsqp_refine_kernel (agents "
                        "re-enqueue themselves after each SQP iteration)"}

    cpu_baseline = cpu_1core = cpu_banded = parity = None
    if not args.no_cpu_baseline and world == 1:   # the CPU legs are reported at N = 1 only
        cores = os.cpu_count() or 1
        sample, chosen = cpu_sample(p, inst, 12.0 * cores * 0.7, cores)
        t, r = cpu_port_run(p, sample, cores)
        cpu_baseline = {"value": int(r.n_qp.sum()) / t, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{len(chosen)} of {n_inst} instances ({sample.n_agents} agents), one pass, {t:.1f} s; "
                                  "KKT LDL^T path, OpenMP over agents",
                        "refine_ms_per_instance": 1e3 * t / len(chosen)}
        # full parity of the CUDA result with the oracle on the sampled instances (same inputs, same planes)
        gres = dr.to_host()
        a_idx = np.concatenate([np.arange(batch.inst_agent_ptr[i], batch.inst_agent_ptr[i + 1]) for i in chosen])
        gt = np.concatenate([gres.agent_traj(batch, a).reshape(-1) for a in a_idx])
        gc = np.concatenate([gres.agent_corridor(batch, a).reshape(-1) for a in a_idx])
        pk = np.concatenate([batch.plane_abc[12 * batch.plane_ptr[a]:12 * batch.plane_ptr[a + 1]] for a in a_idx])
        # per-agent agreement: the SQP loop regenerates corridors in 0.1 m steps and cuts ADMM off at 400
        # iterations, so a 1e-9 difference between two linear solvers occasionally flips a discrete decision
        # and that agent's trajectory then differs visibly (SURVEY finding 3); the counts say how often
        dmax = np.zeros(len(a_idx))
        pos = 0
        for j, a in enumerate(a_idx):
            n6 = 6 * int(batch.agent_off[a + 1] - batch.agent_off[a])
            dmax[j] = np.abs(gt[pos:pos + n6] - r.traj[pos:pos + n6]).max(); pos += n6
        succ_g = np.abs(gres.inst_status[chosen]) <= 2
        succ_o = np.abs(r.inst_status) <= 2
        parity = {"vs": "oracle (CPU port, KKT path)", "instances": len(chosen), "agents": int(sample.n_agents),
                  "planes_bit_equal": bool(pk.shape == sample.plane_abc.shape and np.array_equal(pk, sample.plane_abc)),
                  "status_equal": bool(np.array_equal(gres.status[a_idx], r.status)),
                  "sqp_iters_equal": bool(np.array_equal(gres.sqp_iters[a_idx], r.sqp_iters)),
                  "admm_iters_equal": bool(np.array_equal(gres.admm_iters[a_idx], r.admm_iters)),
                  "n_factor_equal": bool(np.array_equal(gres.n_factor[a_idx], r.n_factor)),
                  "inst_status_equal": bool(np.array_equal(gres.inst_status[chosen], r.inst_status)),
                  "success_equal": bool(np.array_equal(succ_g, succ_o)),
                  "success_rate": {"cuda": float(succ_g.mean()), "oracle": float(succ_o.mean())},
                  "max_abs_traj": float(dmax.max()), "median_abs_traj": float(np.median(dmax)),
                  "max_abs_corridor": float(np.abs(gc - r.corridors).max()),
                  "agents_within_1e-6": int((dmax < 1e-6).sum()), "agents_within_1e-3": int((dmax < 1e-3).sum()),
                  "agents_differing_in_status": int((gres.status[a_idx] != r.status).sum()),
                  "agents_differing_in_admm_iters": int((gres.admm_iters[a_idx] != r.admm_iters).sum())}
        # what ./csdo does: one core, agents one after the other (dsqp_solver.cc:1198); and the banded variant
        one = pack_instances(with_oracle_planes(p, [inst[i] for i in chosen[:max(1, len(chosen) // cores)]]))
        t1, r1 = cpu_port_run(p, one, 1)
        cpu_1core = {"value": int(r1.n_qp.sum()) / t1, "unit": UNIT, "cores": 1, "kind": "port",
                     "sample": f"{one.n_inst} instances ({one.n_agents} agents) sequentially on one core, {t1:.1f} s",
                     "refine_ms_per_instance": 1e3 * t1 / one.n_inst}
        tb, rb = cpu_port_run(p, sample, cores, linsys=1)
        cpu_banded = {"value": int(rb.n_qp.sum()) / tb, "unit": UNIT, "cores": cores, "kind": "port",
                      "sample": f"same sample, reduced banded system instead of the KKT LDL^T (linsys=1), {tb:.1f} s"}
        # the yardstick for the per-agent differences above: the oracle against ITSELF with the other linear
        # solver (KKT LDL^T vs reduced banded system, both CPU, both exact solves differing at rounding level)
        dself = np.zeros(len(a_idx))
        pos = 0
        for j, a in enumerate(a_idx):
            n6 = 6 * int(batch.agent_off[a + 1] - batch.agent_off[a])
            dself[j] = np.abs(rb.traj[pos:pos + n6] - r.traj[pos:pos + n6]).max(); pos += n6
        parity["oracle_kkt_vs_oracle_banded"] = {
            "what": "the same CPU oracle with its two linear-system paths on the same sample: how far rounding-level "
                    "differences alone move the result of this algorithm",
            "status_equal": bool(np.array_equal(rb.status, r.status)),
            "admm_iters_equal": bool(np.array_equal(rb.admm_iters, r.admm_iters)),
            "agents_differing_in_admm_iters": int((rb.admm_iters != r.admm_iters).sum()),
            "max_abs_traj": float(dself.max()), "median_abs_traj": float(np.median(dself)),
            "agents_within_1e-6": int((dself < 1e-6).sum()), "agents_within_1e-3": int((dself < 1e-3).sum())}
    latency = None
    if world == 1 and not args.no_cpu_baseline:
        latency = latency_block(p, DsqpSolver, local_rank)

    par = (f"agents of every instance partitioned x{world}, one NCCL all-gather of trajectories/statuses per step"
           if agents_mode else f"instance-sharded x{world} (rank r refines instances r, r+{world}, ...), no data-path collective")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "real benchmark scenarios, stand-in coarse plans" if args.workload == "real" else "synthetic",
            "refine_ms_per_instance": 1e3 * t_dev / args.steps / max(tot_inst, 1),   # whole job: step time / all instances
            "config": config_dict(args, total),
            "details": {"instances_total": tot_inst, "agents_total": tot_agents, "qp_per_step_total": tot_qp // args.steps,
                        "admm_iters_per_step_total": admm_tot, "planes_rank0": int(batch.plane_ptr[-1]),
                        "agent_steps_rank0": int(batch.agent_nt()[own].sum()), "horizon_max": int(batch.inst_nt.max()),
                        "parallelism": par, "launch": launch, "generation_s": t_gen},
            "clocks": clk, "gpu_launches": int(launch.get("launches", 0)) * args.steps,
            "e2e": {"value": tot_qp / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * t_e2e / args.steps, "steps_timed": e2e_steps, "matches_device_arm": e2e_ok,
                    "timing": "wall clock around the blocking csdo_refine() calls (pinned host buffers)"},
            "preprocess": {"ms": float(np.median(pre_ms)), "what": "csdo_planes_count_device + csdo_planes_fill_device "
                           "(neighbour pairs + planes, device-resident, rank 0's share); the reference's rt_preprocess",
                           "planes": int(batch.plane_ptr[-1])},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "cpu_baseline_1core": cpu_1core,
            "cpu_baseline_banded": cpu_banded, "parity": parity, "latency": latency}
    if gather is not None:
        line["allgather"] = {"ms_per_step": float(np.mean(gather_ms)), "share_of_step": float(np.mean(gather_ms) / np.mean(step_ms)),
                             "bytes_per_rank": gather.bytes_per_rank, "bit_identical_to_unsharded": bit_identical}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
