#!/bin/bash
# next-head key requested together with the record (one L2 round trip per R_ARRIVE / R_BUFFER instead of two): parity + timing
cd "$(dirname "$0")/.."
cat > /tmp/b23.py <<PY
import os, sys, subprocess, tempfile
sys.path.insert(0, os.getcwd())
from tools import gpu_probe
conns = sys.argv[1]
d = tempfile.mkdtemp()
subprocess.run([sys.executable, "tools/gen_topology.py", "64", conns, d, "--name", "big", "--routing", "prog-adaptive"], check=True, stdout=subprocess.DEVNULL)
gpu_probe.perf(os.path.join(d, "big.conf"), reps=3, num_messages=10, traffic=1)
PY
python tools/gpu_probe.py parity perf 2>&1 | grep "IDENT\|DIFF\|ERROR\|Mev/s" | cut -c1-150
python /tmp/b23.py 16 2>&1 | grep "Mev/s" | sed "s/^/[1056 routers] /" | cut -c1-150
python /tmp/b23.py 8 2>&1 | grep "Mev/s" | sed "s/^/[2080 routers] /" | cut -c1-150
python /tmp/b23.py 2 2>&1 | grep "Mev/s" | sed "s/^/[8224 routers] /" | cut -c1-150
