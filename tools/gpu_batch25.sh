#!/bin/bash
# radix-scaled activation budget as the default: check on the partition sizes of 4- and 8-GPU runs, the BASELINE configs, and the hysteresis knob
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
python /tmp/b23.py 16 2>&1 | grep "Mev/s" | sed "s/^/[1056 routers, default] /" | cut -c1-150
python /tmp/b23.py 8 2>&1 | grep "Mev/s" | sed "s/^/[2080 routers, default] /" | cut -c1-150
for h in 0.1 0.5 0.8; do
DFD_HYST_FRAC=$h python /tmp/b23.py 16 2>&1 | grep "Mev/s" | sed "s/^/[1056 routers, hyst $h] /" | cut -c1-150
done
DFD_IDLE_SLEEP_NS=500 python /tmp/b23.py 16 2>&1 | grep "Mev/s" | sed "s/^/[1056 routers, idle sleep 500] /" | cut -c1-150
python tools/gpu_probe.py perf 2>&1 | grep "Mev/s" | cut -c1-150
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
