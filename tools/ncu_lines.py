"""Attribute the warp-stall samples of an ncu report to source lines.
usage: python tools/ncu_lines.py <report.ncu-rep> <cubin> <kernel-substring> [top N]
Joins `ncu --page source --csv` (per SASS instruction, in program order) with `nvdisasm -g` line info of the same cubin."""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

rep, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
cols = rows[hdr]
ci = {c: i for i, c in enumerate(cols)}
inst = [r for r in rows[hdr + 1:] if len(r) == len(cols)]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
lines, cur, active = [], None, False
for l in dis:
    if l.startswith(".text."):
        active = kname in l
        continue
    if not active:
        continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
if len(lines) != len(inst):
    print(f"warning: {len(lines)} instructions in the cubin vs {len(inst)} in the report", file=sys.stderr)
n = min(len(lines), len(inst))
stall_cols = [c for c in cols if c.startswith("stall_") and "Not Issued" not in c]
per_line = defaultdict(lambda: defaultdict(int))
total = 0
for k in range(n):
    s = int(inst[k][ci["# Samples"]] or 0)
    total += s
    if not s:
        continue
    per_line[lines[k]]["samples"] += s
    per_line[lines[k]]["inst"] += int(inst[k][ci["Instructions Executed"]] or 0)
    for c in stall_cols:
        v = int(inst[k][ci[c]] or 0)
        if v:
            per_line[lines[k]][c] += v
print(f"total samples {total}")
agg = defaultdict(int)
for ln, d in per_line.items():
    for c in stall_cols:
        agg[c] += d.get(c, 0)
print("stall totals:", ", ".join(f"{c[6:]} {100 * v / total:.1f}%" for c, v in sorted(agg.items(), key=lambda x: -x[1]) if v * 100 > total))
for ln, d in sorted(per_line.items(), key=lambda x: -x[1]["samples"])[:top]:
    reasons = sorted(((c[6:], d[c]) for c in stall_cols if d.get(c)), key=lambda x: -x[1])[:3]
    print(f"{100 * d['samples'] / total:5.1f}%  {ln[0]}:{ln[1]:<5d} " + " ".join(f"{r}={v}" for r, v in reasons))
