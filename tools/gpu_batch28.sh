#!/bin/bash
# 4,128 routers (129 groups: what a GPU holds at N=2): pooled kernel against the static one-warp-per-block kernel, budgets
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
python /tmp/b23.py 4 2>&1 | grep "Mev/s" | sed "s/^/[4128 routers, default] /" | cut -c1-170
DFD_POOL=0 python /tmp/b23.py 4 2>&1 | grep "Mev/s" | sed "s/^/[4128 routers, static] /" | cut -c1-170
DFD_POOL=0 DFD_ACT_BUDGET=10 python /tmp/b23.py 4 2>&1 | grep "Mev/s" | sed "s/^/[4128 routers, static, budget 10] /" | cut -c1-170
DFD_ACT_BUDGET=10 python /tmp/b23.py 4 2>&1 | grep "Mev/s" | sed "s/^/[4128 routers, pooled, budget 10] /" | cut -c1-170
DFD_ACT_BUDGET=24 python /tmp/b23.py 4 2>&1 | grep "Mev/s" | sed "s/^/[4128 routers, pooled, budget 24] /" | cut -c1-170
DFD_ACT_BUDGET=10 python /tmp/b23.py 2 2>&1 | grep "Mev/s" | sed "s/^/[8224 routers, pooled, budget 10] /" | cut -c1-170
DFD_ACT_BUDGET=24 python /tmp/b23.py 2 2>&1 | grep "Mev/s" | sed "s/^/[8224 routers, pooled, budget 24] /" | cut -c1-170
DFD_IDLE_SLEEP_NS=2000 python /tmp/b23.py 2 2>&1 | grep "Mev/s" | sed "s/^/[8224 routers, pooled, idle sleep 2000] /" | cut -c1-170
DFD_IDLE_SLEEP_NS=0 python /tmp/b23.py 2 2>&1 | grep "Mev/s" | sed "s/^/[8224 routers, pooled, idle sleep 0] /" | cut -c1-170
