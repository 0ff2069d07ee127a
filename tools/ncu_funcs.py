"""Per-function share of the warp-stall samples of an ncu report (dfd_core.h functions by line range).
usage: python tools/ncu_funcs.py <report.ncu-rep> <cubin> <kernel-substring>"""
import bisect
import os
import re
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(ROOT, "codes_b200/csrc/dfd_core.h")).read().splitlines()
funcs = []
for i, l in enumerate(src, 1):
    m = re.match(r"^(?:template <[^>]*> )?DFD_FN(?:_NOINLINE)? [\w:<>\*& ]+?[\s\*&](\w+)\(", l)
    if m:
        funcs.append((i, m.group(1)))
starts = [f[0] for f in funcs]
out = subprocess.run([sys.executable, os.path.join(ROOT, "tools/ncu_lines.py"), sys.argv[1], sys.argv[2], sys.argv[3], "100000"], capture_output=True, text=True).stdout.splitlines()
print(out[0])
print(out[1][:400])
agg = defaultdict(float)
reasons = defaultdict(lambda: defaultdict(int))
for l in out[2:]:
    m = re.match(r"\s*([\d.]+)%\s+(\S+):(\d+)\s*(.*)", l)
    if not m:
        continue
    pct, f, ln = float(m.group(1)), m.group(2), int(m.group(3))
    if f == "dfd_core.h":
        k = bisect.bisect_right(starts, ln) - 1
        name = funcs[k][1] if k >= 0 else "?"
    else:
        name = f
    agg[name] += pct
for k, v in sorted(agg.items(), key=lambda x: -x[1])[:45]:
    print(f"{v:6.1f}% {k}")
