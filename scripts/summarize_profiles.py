"""Developer tooling: turn gpurun_out/ ncu outputs into the tracked summaries under profiles/.

usage: python scripts/summarize_profiles.py <round tag, e.g. r02> <workload, e.g. c5>
  gpurun_out/<tag>_launches.csv         (ncu --metrics gpu__time_duration.sum ... python bench.py ...)
  gpurun_out/<tag>_refine_full.ncu-rep  (ncu --set full -k regex:dsqp_refine -c 1 ... python bench.py ...)
  gpurun_out/<tag>_bench_under_ncu.log  (stdout of that bench run: its JSON line gives the agent steps)
"""
import collections, csv, io, json, os, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
wl = sys.argv[2] if len(sys.argv) > 2 else "c5"
rows = list(csv.reader(open(f"gpurun_out/{tag}_launches.csv")))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in data:
    name = r[kn].split("(")[0]
    v = float(r[mv].replace(",", "")); u = r[mu]
    ms = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}[u] * v
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ms
tot = sum(a[1] for a in agg.values())
lines = [f"# {tag} ncu launch list: python bench.py --workload {wl} ... (gpu__time_duration.sum, --clock-control none)",
         "# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes",
         "kernel,launches,total_ms,share"]
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{k},{n},{ms:.3f},{ms / tot:.4f}")
open(f"profiles/{tag}_launch_list.csv", "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:8]))
out = subprocess.run(["ncu", "-i", f"gpurun_out/{tag}_refine_full.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(out)))
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "lts__t_sector_hit_rate.pct"]
summ = {a: (c, b) for a, b, c in zip(rr[0], rr[1], rr[2]) if a in want}
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tr = sum(float(summ[k][0].replace(",", "")) * mult[summ[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
line = [l for l in open(f"gpurun_out/{tag}_bench_under_ncu.log") if l.startswith("{")][-1]
steps = json.loads(line)["details"]["agent_steps_rank0"]
git = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
tj = {}
if os.path.exists("profiles/traffic.json"):
    tj = json.load(open("profiles/traffic.json"))
tj = {k: v for k, v in tj.items() if k.startswith("dram_bytes_per_agent_step_")}
tj[f"dram_bytes_per_agent_step_{wl}"] = tr / steps
tj["source"] = f"profiles/{tag}_refine_full_summary.txt (ncu --set full, one launch of a {wl} sub-batch, {steps} agent steps)"
tj["git"] = git
json.dump(tj, open("profiles/traffic.json", "w"), indent=1)
st = {a[len("smsp__pcsamp_warps_issue_stalled_"):]: float(c.replace(",", "")) for a, b, c in zip(rr[0], rr[1], rr[2])
      if a.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in a}
stot = sum(st.values()) or 1.0
stalls = "\n".join(f"warp_samples_{k} = {100 * v / stot:.1f} %" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8])
extra = ["lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
         "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
         "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum"]
more = "\n".join(f"{a} = {c} {b}" for a, b, c in zip(rr[0], rr[1], rr[2]) if a in extra)
open(f"profiles/{tag}_refine_full_summary.txt", "w").write(
    f"# ncu --set full --clock-control none -k regex:dsqp_refine -c 1 python bench.py --workload {wl} --steps 1 --no-cpu-baseline (git {git})\n"
    + "\n".join(f"{k} = {summ[k][0]} {summ[k][1]}" for k in want if k in summ) + f"\ndram_bytes_per_launch = {tr:.0f}\nagent_steps_per_launch = {steps}\n" + more + "\n# warp state samples (pc sampling), share of all samples\n" + stalls + "\n")
print(open(f"profiles/{tag}_refine_full_summary.txt").read())
