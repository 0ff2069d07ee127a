#!/bin/bash
# partitions of the size a GPU holds in a 4- and 8-GPU run of the bench network, on one GPU: which kernel / budget is best there?
cd "$(dirname "$0")/.."
cat > /tmp/b23.py <<'PY'
import os, sys, subprocess, tempfile
sys.path.insert(0, os.getcwd())
from tools import gpu_probe
conns = sys.argv[1]
d = tempfile.mkdtemp()
subprocess.run([sys.executable, "tools/gen_topology.py", "64", conns, d, "--name", "big", "--routing", "prog-adaptive"], check=True, stdout=subprocess.DEVNULL)
gpu_probe.perf(os.path.join(d, "big.conf"), reps=3, num_messages=10, traffic=1)
PY
echo "--- 2,080 routers (65 groups): what a GPU holds at N=4"
python /tmp/b23.py 8 2>&1 | grep "Mev/s" | sed "s/^/[default] /" | cut -c1-300
DFD_WIDE_REGS=1 python /tmp/b23.py 8 2>&1 | grep "Mev/s" | sed "s/^/[wide regs] /" | cut -c1-300
DFD_POOL=1 python /tmp/b23.py 8 2>&1 | grep "Mev/s" | sed "s/^/[pool] /" | cut -c1-300
DFD_ACT_BUDGET=8 python /tmp/b23.py 8 2>&1 | grep "Mev/s" | sed "s/^/[budget 8] /" | cut -c1-300
echo "--- 1,056 routers (33 groups): what a GPU holds at N=8"
python /tmp/b23.py 16 2>&1 | grep "Mev/s" | sed "s/^/[default] /" | cut -c1-300
DFD_ACT_BUDGET=2 python /tmp/b23.py 16 2>&1 | grep "Mev/s" | sed "s/^/[budget 2] /" | cut -c1-300
DFD_ACT_BUDGET=8 python /tmp/b23.py 16 2>&1 | grep "Mev/s" | sed "s/^/[budget 8] /" | cut -c1-300
DFD_ACT_BUDGET=16 python /tmp/b23.py 16 2>&1 | grep "Mev/s" | sed "s/^/[budget 16] /" | cut -c1-300
DFD_POOL=1 python /tmp/b23.py 16 2>&1 | grep "Mev/s" | sed "s/^/[pool] /" | cut -c1-300
DFD_PROFILE_PRINT=1 python /tmp/b23.py 16 2>&1 | grep "cycles\|rounds\|warps" | tail -4 | cut -c1-300
