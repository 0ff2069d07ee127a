"""Executed warp instructions per source line (top N) of an ncu report.  usage: ncu_lines_inst.py <rep> <cubin> <kernel-substr> [events] [top]"""
import csv, io, re, subprocess, sys
from collections import defaultdict
rep, cubin, kname = sys.argv[1:4]
events = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
top = int(sys.argv[5]) if len(sys.argv) > 5 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
cols = rows[hdr]; ci = {c: i for i, c in enumerate(cols)}
inst = [r for r in rows[hdr + 1:] if len(r) == len(cols)]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
lines, cur, active = [], None, False
for l in dis:
    if l.startswith(".text."):
        active = kname in l; continue
    if not active: continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l): lines.append(cur)
n = min(len(lines), len(inst))
E = defaultdict(int); S = defaultdict(int); tot = 0
for k in range(n):
    e = int(inst[k][ci["Instructions Executed"]] or 0)
    E[lines[k]] += e; S[lines[k]] += 1; tot += e
print(f"total {tot/events:.0f} inst/event")
for ln, e in sorted(E.items(), key=lambda x: -x[1])[:top]:
    print(f"{e/events:7.1f}  {ln[0] if ln else '?'}:{ln[1] if ln else 0:<5d} sass {S[ln]}")
