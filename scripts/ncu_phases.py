"""Developer tooling: aggregate ncu per-instruction samples by the call-site line inside a given
function (inline chains from `nvdisasm -gi`), i.e. per phase of solve_qp / refine_body.

usage: ncu_phases.py <report.ncu-rep> <cubin> <kernel-substring> <lo> <hi> [phase line to break down by innermost source line]
Every SASS instruction is attributed to the outermost frame of its inline chain whose line lies in
[lo, hi] of dsqp_kernel.cu; instructions without such a frame go to "other".
"""
import csv, io, re, subprocess, sys
from collections import defaultdict

rep, cubin, kname, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
fsub = "dsqp_kernel.cu"
inner = int(sys.argv[6]) if len(sys.argv) > 6 else None  # optional: break this phase line down by innermost line
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ci = {n: hdr.index(n) for n in ("Source", "# Samples", "Instructions Executed")}
stall_cols = [(n, hdr.index(n)) for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
inst = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
tags, chain, infn, fresh = [], [], False, True
for ln in dis:
    if ln.startswith(".text."):
        infn = kname in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if fresh:
            chain = []
            fresh = False
        chain.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        fresh = True
        tag = "other"
        for f, l in reversed(chain):  # outermost frame first
            if fsub in f and lo <= l <= hi:
                tag = l
                break
        tags.append(tag if inner is None or tag != inner else (tag, chain[0]))
print("ncu instructions:", len(inst), "nvdisasm instructions:", len(tags))
agg = defaultdict(lambda: [0, 0, 0, defaultdict(int)])
tot = 0
for r, l in zip(inst, tags):
    s = int(float(r[ci["# Samples"]] or 0)); e = int(float(r[ci["Instructions Executed"]] or 0))
    a = agg[l]; a[0] += s; a[1] += e; a[2] += 1; tot += s
    for n, i in stall_cols:
        v = r[i]
        if v and v != "0":
            a[3][n] += int(float(v))
print("total samples", tot)
for l, (s, e, n, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:60]:
    tops = ", ".join(f"{k[6:]}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:5])
    print(f"{100.0 * s / max(tot, 1):6.2f}%  warp-inst {e:>12}  sass {n:>6}  line {l}  [{tops}]")
