"""Developer tooling: aggregate ncu per-instruction stall samples by CUDA source line.

usage: ncu_lines.py <report.ncu-rep> <cubin> <kernel-substring> [top]
Joins `ncu --page source --csv` (SASS order) with `nvdisasm -g` (line info) by instruction index.
"""
import csv, io, re, subprocess, sys
from collections import defaultdict

rep, cubin, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ci = {n: hdr.index(n) for n in ("Source", "# Samples", "Instructions Executed")}
stall_cols = [(n, hdr.index(n)) for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
inst = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
lines, cur, infn = [], None, False
for ln in dis:
    if ln.startswith(".text."):
        infn = kname in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
print("ncu instructions:", len(inst), "nvdisasm instructions:", len(lines))
agg = defaultdict(lambda: [0, 0, defaultdict(int)])
tot = 0
for r, l in zip(inst, lines):
    s = int(float(r[ci["# Samples"]] or 0)); e = int(float(r[ci["Instructions Executed"]] or 0))
    a = agg[l]; a[0] += s; a[1] += e; tot += s
    for n, i in stall_cols:
        v = r[i]
        if v and v != "0":
            a[2][n] += int(float(v))
print("total samples", tot)
for l, (s, e, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    tops = ", ".join(f"{n[6:]}:{v}" for n, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100.0 * s / max(tot, 1):6.2f}%  inst {e:>12}  {l}  [{tops}]")
