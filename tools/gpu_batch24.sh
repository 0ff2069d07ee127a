#!/bin/bash
# event budget of an activation in the one-warp-per-group kernels, by network shape (the default 4 was tuned on radix 31)
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
for b in 4 6 8 10 12; do
DFD_ACT_BUDGET=$b python /tmp/b23.py 16 2>&1 | grep "Mev/s" | sed "s/^/[radix 63, 1056 routers, budget $b] /" | cut -c1-160
done
for b in 6 8 10 12; do
DFD_ACT_BUDGET=$b python /tmp/b23.py 8 2>&1 | grep "Mev/s" | sed "s/^/[radix 63, 2080 routers, budget $b] /" | cut -c1-160
done
for b in 2 3 4 6 8; do
DFD_ACT_BUDGET=$b python tools/gpu_probe.py perf 2>&1 | grep "Mev/s" | grep -v "dfdally-72" | sed "s/^/[budget $b] /" | cut -c1-160
done
