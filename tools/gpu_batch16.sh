#!/bin/bash
# informational: bench network (131,584 terminals, PAR, uniform) with more messages per terminal; 1,050,624 terminals on the current build
cd "$(dirname "$0")/.."
python - <<'PY' 2>&1 | grep -v "^$" | cut -c1-400
import os, sys, subprocess, tempfile
sys.path.insert(0, os.getcwd())
from tools import gpu_probe
d = tempfile.mkdtemp()
subprocess.run([sys.executable, "tools/gen_topology.py", "64", "2", d, "--name", "big", "--routing", "prog-adaptive"], check=True, stdout=subprocess.DEVNULL)
for m in (10, 20, 50):
    gpu_probe.perf(os.path.join(d, "big.conf"), reps=2, num_messages=m, traffic=1)
gpu_probe.perf(os.path.join(d, "big.conf"), reps=2, num_messages=50, traffic=3)
PY
python tools/gpu_probe.py big1m 2>&1 | grep "Mev/s" | cut -c1-400
DFD_PREFETCH=1 python tools/gpu_probe.py big1m 2>&1 | grep "Mev/s" | sed 's/^/[prefetch] /' | cut -c1-400
