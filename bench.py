#!/usr/bin/env python
"""Benchmark of the DSQP refine path (BASELINE.json: batched agent-QP solves/sec; refine ms/instance).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one refine pass of the hot path over one batch of synthetic instances.  Workload at any N:
per GPU the shape of BASELINE.json configs[1] -- the full map50by50 sweep: agents 5/10/15/20/25 x
{empty, 25 obstacles} x 60 instances = 600 instances / 9000 agents, seeded synthetic priority-style
plans (the PBS + Hybrid-A* front end is out of scope and cannot be built offline).  Weak scaling: every
rank refines its own 600 instances (instance sharding, no data-path collective).

`value`  : device-resident inputs (torch tensors in HBM), CUDA events on the launching stream.
`e2e`    : the same metric through csdo_refine() with pinned HOST buffers, H2D + D2H inside the timed region.
`--impl reference`: the reference's CPU path.  The reference does not compile offline (Eigen, OSQP 0.6.3,
yaml-cpp, Boost, OMPL absent), so this arm times the CPU oracle port (oracle/, OpenMP over agents, all
host threads) on a bounded sample of the same workload.
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
PER_SHAPE = int(os.environ.get("CSDO_BENCH_PER_SHAPE", "60"))   # 60 = the benchmark's instances per folder


def build_workload(rank: int):
    """600 synthetic instances of the map50by50 sweep shape for this rank (seeded)."""
    from csdotrajectoryplanning_b200 import default_params
    from csdotrajectoryplanning_b200.scenario import MAP50_SWEEP, synthetic_batch
    p = default_params()
    inst = synthetic_batch(MAP50_SWEEP, PER_SHAPE, seed=1234 + 100000 * rank, params=p)
    return p, inst


def algorithmic_flops(batch, res) -> float:
    """F_QP = n_fac*504*Nt + n_it*(518*Nt + 88*K) + floor(n_it/25)*F_chk + F_scale,
    evaluated with the COUNTED iterations/factorizations of every agent (summed over its QPs)."""
    nt = batch.agent_nt().astype(np.float64)
    K = np.diff(batch.plane_ptr).astype(np.float64)
    nnzA, nnzP, n, m = 28 * nt - 11 + 12 * K, 5 * nt, 6 * nt - 2, 13 * nt + 4 * K
    f_chk = 6 * nnzA + 2 * nnzP + 8 * (n + m)
    f_scale = 10 * 4 * (nnzA + nnzP)
    it, fac, nqp = res.admm_iters.astype(np.float64), res.n_factor.astype(np.float64), res.n_qp.astype(np.float64)
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


def cpu_port_run(p, batch, nthreads: int):
    """The CPU oracle port on `batch` (KKT LDL^T path = the cost profile of OSQP/QDLDL)."""
    from oracle import oracle as O   # bench.py's cpu_baseline / --impl reference legs only
    t0 = time.perf_counter()
    res, _ = O.refine(p, batch, linsys=0, nthreads=nthreads)
    return time.perf_counter() - t0, res


def cpu_sample(p, inst, target_s: float, nthreads: int):
    """Bounded sample: the first instances of the workload worth about target_s of CPU time."""
    from csdotrajectoryplanning_b200 import pack_instances
    probe = pack_instances(inst[:: max(1, len(inst) // 8)][:8])
    t, r = cpu_port_run(p, probe, nthreads)
    per_agent = t / max(1, probe.n_agents)
    n_agents_target = max(probe.n_agents, int(target_s / max(per_agent, 1e-9)))
    chosen, tot = [], 0
    stride = max(1, len(inst) // 60)
    first = list(range(0, len(inst), stride))             # spread over the shapes first
    for i in first + [i for i in range(len(inst)) if i % stride]:
        if tot >= n_agents_target:
            break
        chosen.append(inst[i]); tot += inst[i].n_agents
    return pack_instances(chosen), len(chosen)


def attach_planes_gpu(solver, inst):
    from csdotrajectoryplanning_b200 import pack_instances
    b0 = pack_instances(inst)
    pb, _ = solver.planes(b0)
    return pb


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    p, inst = build_workload(0)
    from csdotrajectoryplanning_b200 import pack_instances
    from oracle import oracle as O
    for ins in inst:
        ins.plane_t, ins.plane_abc, _ = O.instance_planes(p, ins.guess)
    cores = os.cpu_count() or 1
    budget = 150.0 / max(1, args.steps + args.warmup)          # whole run within a few minutes
    sample, n_inst = cpu_sample(p, inst, min(20.0, budget) * cores * 0.7, cores)
    for _ in range(args.warmup):
        cpu_port_run(p, sample, cores)
    t_tot, qps = 0.0, 0
    for _ in range(args.steps):
        t, r = cpu_port_run(p, sample, cores)
        t_tot += t; qps += int(r.n_qp.sum())
    value = qps / t_tot
    desc = f"{n_inst} of 600 instances ({sample.n_agents} agents) of the map50by50-sweep-shaped workload per step"
    line = {"metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "refine_ms_per_instance": 1e3 * t_tot / args.steps / n_inst,
            "config": {"workload": "map50by50 sweep shape (agents 5-25 x empty/obstacle), synthetic priority-style plans",
                       "sample": desc},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference does not build offline (Eigen/OSQP 0.6.3/yaml-cpp/Boost/OMPL absent): this is the CPU "
                    "oracle port of its OSQP path, OpenMP over agents on all host threads"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
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
    from csdotrajectoryplanning_b200 import binding
    from csdotrajectoryplanning_b200.batch import Batch, RefineResult
    from csdotrajectoryplanning_b200.solver import DeviceBatch, DeviceResult, DsqpSolver

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    p, inst = build_workload(rank)
    solver = DsqpSolver(p, device=local_rank)     # raises without a B200: no CPU fallback
    batch = attach_planes_gpu(solver, inst)       # planes kernels (pre-process, untimed)
    n_inst = batch.n_inst

    # ---- device-resident arm ----
    db, dr = DeviceBatch(batch, dev), DeviceResult(batch, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    # a dedicated non-default stream: the kernels and the timing events share it (stream handle 0
    # would mean "the handle's own stream" to csdo_refine_device)
    stream = torch.cuda.Stream(device=dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]

    def one_step(e=None):
        with torch.cuda.stream(stream):
            flush.fill_(1)                  # L2 flush between iterations (outside the events)
            if e: e[0].record(stream)
            solver.refine_device(db, dr, stream.cuda_stream)
            if e: e[1].record(stream)

    for _ in range(args.warmup):
        one_step()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    torch.cuda.synchronize()
    for k in range(args.steps):
        one_step(ev[k])
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    clk = clocks.stop()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    t_dev = sum(step_ms) * 1e-3
    res = dr.to_host()
    qps_step = int(res.n_qp.sum())
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
    hres = RefineResult.allocate(batch)
    for name in ("traj", "corridors"):
        t, v = pinned_like(getattr(hres, name)); keep.append(t); setattr(hres, name, v)
    h2d = sum(int(v.nbytes) for v in hb.values()) + 4 * batch.n_agents
    d2h = sum(int(getattr(hres, n).nbytes) for n in ("traj", "corridors", "status", "sqp_iters", "n_qp", "admm_iters",
                                                    "n_factor", "objective", "inst_status", "inst_static_legal"))
    for _ in range(2):
        solver.refine(hbatch, hres)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        solver.refine(hbatch, hres)         # blocks until the results are back in host memory
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    e2e_ok = bool(np.array_equal(hres.status, res.status) and np.array_equal(hres.traj, res.traj))

    # ---- reduce over ranks: max time, summed work ----
    tot_qp, tot_inst, tot_agents = qps_step * args.steps, n_inst, batch.n_agents
    if world > 1:
        tt = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_e2e = float(tt[0]), float(tt[1])
        cc = torch.tensor([tot_qp, tot_inst, tot_agents], dtype=torch.float64, device=dev)
        dist.all_reduce(cc, op=dist.ReduceOp.SUM)
        tot_qp, tot_inst, tot_agents = int(cc[0]), int(cc[1]), int(cc[2])
    if rank != 0:
        if world > 1: dist.destroy_process_group()
        return

    value = tot_qp / t_dev
    # ---- roofline of the dominant kernel (dsqp_refine_kernel): FP64 pipe, measured live ----
    fl = algorithmic_flops(batch, res)
    by = algorithmic_bytes(batch)
    kern_s = float(np.mean(step_ms)) * 1e-3
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    import ctypes as C
    fp64 = C.c_double(0.0)
    binding.lib().csdo_measure_fp64_peak(local_rank, C.byref(fp64))
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("dram_bytes_per_launch")
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    roofline = {"bound": "fp64", "achieved": fl / kern_s * 1e-12, "peak": fp64.value, "unit": "TFLOP/s",
                "frac": (fl / kern_s * 1e-12) / fp64.value if fp64.value else None, "traffic": traffic,
                "peak_source": "csdo_measure_fp64_peak (DFMA microbenchmark, this run); FP64 is not in MEASURED_PEAKS.json",
                "algorithmic_flops_per_launch": fl,
                "hbm": {"bound": "hbm", "achieved": by / kern_s * 1e-9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": by / kern_s * 1e-9 / hbm_peak, "algorithmic_bytes_per_launch": by,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)"},
                "launches_per_step": 1,
                "note": "banded FP64 ADMM with the working set in shared memory: neither HBM nor tensor cores bind it; "
                        "one step = one persistent launch of dsqp_refine_kernel (agents re-enqueue themselves after "
                        "each SQP iteration)"}

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:   # the CPU baseline is reported at N = 1 only
        inst = batch.unpack()               # instances with the planes built above
        cores = os.cpu_count() or 1
        sample, ns = cpu_sample(p, inst, 12.0 * cores * 0.7, cores)
        t, r = cpu_port_run(p, sample, cores)
        cpu_baseline = {"value": int(r.n_qp.sum()) / t, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{ns} of 600 instances ({sample.n_agents} agents), one pass, {t:.1f} s",
                        "refine_ms_per_instance": 1e3 * t / ns}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "refine_ms_per_instance": 1e3 * t_dev / args.steps / tot_inst,   # whole job: step time / all instances
            "config": {"workload": "map50by50 full sweep shape: agents 5/10/15/20/25 x {empty, 25 obstacles} x "
                                   f"{PER_SHAPE} = {n_inst} instances ({batch.n_agents} agents) per GPU, synthetic "
                                   "priority-style plans, one batch per GPU",
                       "instances_total": tot_inst, "agents_total": tot_agents, "qp_per_step_total": tot_qp // args.steps,
                       "admm_iters_per_step_rank0": int(res.admm_iters.sum()),
                       "horizon_max": int(batch.inst_nt.max()), "l2": "256 MiB buffer written between timed iterations",
                       "parallelism": f"instance-sharded x{world}, no data-path collective", "launch": launch},
            "clocks": clk, "gpu_launches": int(launch.get("launches", 0)) * args.steps,
            "e2e": {"value": tot_qp / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * t_e2e / args.steps, "matches_device_arm": e2e_ok,
                    "timing": "wall clock around the blocking csdo_refine() calls (pinned host buffers)"},
            "roofline": roofline, "cpu_baseline": cpu_baseline}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
