"""Per-function share of stall samples AND executed warp instructions of an ncu report (dfd_core.h functions by line range).
usage: python tools/ncu_funcs2.py <report.ncu-rep> <cubin> <kernel-substring> [events]"""
import bisect, csv, io, os, re, subprocess, sys
from collections import defaultdict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, cubin, kname = sys.argv[1:4]
events = float(sys.argv[4]) if len(sys.argv) > 4 else None
src = open(os.path.join(ROOT, "codes_b200/csrc/dfd_core.h")).read().splitlines()
funcs = []
for i, l in enumerate(src, 1):
    m = re.match(r"^(?:template <[^>]*> )?DFD_FN(?:_NOINLINE)? [\w:<>\*& ]+?[\s\*&](\w+)\(", l)
    if m:
        funcs.append((i, m.group(1)))
starts = [f[0] for f in funcs]
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
n = min(len(lines), len(inst))
S, I, N = defaultdict(int), defaultdict(int), defaultdict(int)
ts = ti = 0
for k in range(n):
    s = int(inst[k][ci["# Samples"]] or 0)
    e = int(inst[k][ci["Instructions Executed"]] or 0)
    f, ln = lines[k] if lines[k] else ("?", 0)
    if f == "dfd_core.h":
        j = bisect.bisect_right(starts, ln) - 1
        name = funcs[j][1] if j >= 0 else "?"
    else:
        name = f
    S[name] += s; I[name] += e; N[name] += 1
    ts += s; ti += e
print(f"samples {ts}, warp instructions {ti}" + (f", {ti/events:.0f} per event" if events else ""))
print(f"{'function':28s} {'samples%':>8s} {'inst%':>7s} {'SASS':>6s}" + ("  inst/event" if events else ""))
for k, v in sorted(S.items(), key=lambda x: -x[1])[:50]:
    print(f"{k:28s} {100*v/ts:8.1f} {100*I[k]/ti:7.1f} {N[k]:6d}" + (f"  {I[k]/events:8.1f}" if events else ""))
