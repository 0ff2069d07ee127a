"""Warp instructions executed per source function of dfd_core.h (from an ncu report with SourceCounters).
usage: python tools/ncu_inst.py <report.ncu-rep> <cubin> <kernel-substring> <events>"""
import bisect
import csv
import io
import os
import re
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, cubin, kname, events = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
src = open(os.path.join(ROOT, "codes_b200/csrc/dfd_core.h")).read().splitlines()
funcs = []
for i, l in enumerate(src, 1):
    m = re.match(r"^(?:template <[^>]*> )?DFD_FN(?:_NOINLINE)? [\w:<>\*& ]+?[\s\*&](\w+)\(", l)
    if m:
        funcs.append((i, m.group(1)))
starts = [f[0] for f in funcs]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
cols = rows[h]
ci = {c: i for i, c in enumerate(cols)}
inst = [r for r in rows[h + 1:] if len(r) == len(cols)]
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
agg = defaultdict(lambda: [0, 0, 0])
for k in range(min(len(lines), len(inst))):
    n = int(inst[k][ci["Instructions Executed"]] or 0)
    s = int(inst[k][ci["# Samples"]] or 0)
    f, ln = lines[k]
    if f == "dfd_core.h":
        j = bisect.bisect_right(starts, ln) - 1
        name = funcs[j][1] if j >= 0 else "?"
    else:
        name = f
    agg[name][0] += n
    agg[name][1] += s
    agg[name][2] += 1
tot = sum(v[0] for v in agg.values())
tots = sum(v[1] for v in agg.values())
print(f"warp instructions {tot} ({tot / events:.0f} per event), samples {tots}")
for k, v in sorted(agg.items(), key=lambda x: -x[1][0])[:40]:
    print(f"{v[0] / events:8.1f} inst/event {100 * v[0] / tot:5.1f}%  samples {100 * v[1] / tots:5.1f}%  cyc/inst {v[1] / max(v[0], 1) * tot / tots:5.2f}x  sass {v[2]:5d}  {k}")
