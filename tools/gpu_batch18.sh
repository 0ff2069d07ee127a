#!/bin/bash
# poll backoff in the pooled kernel (needs experiments/poll_backoff.patch applied): bench network, 10 and 50 messages per terminal
cd "$(dirname "$0")/.."
cat > /tmp/b18.py <<'PY'
import os, sys, subprocess, tempfile
sys.path.insert(0, os.getcwd())
from tools import gpu_probe
d = tempfile.mkdtemp()
subprocess.run([sys.executable, "tools/gen_topology.py", "64", "2", d, "--name", "big", "--routing", "prog-adaptive"], check=True, stdout=subprocess.DEVNULL)
for m in [int(x) for x in sys.argv[1:]]:
    gpu_probe.perf(os.path.join(d, "big.conf"), reps=3, num_messages=m, traffic=1)
PY
for b in 0 1 2 3 5; do
DFD_POLL_BACKOFF=$b python /tmp/b18.py 10 2>&1 | grep "Mev/s" | sed "s/^/[backoff $b] /" | cut -c1-330
done
for b in 0 2 4; do
DFD_POLL_BACKOFF=$b python /tmp/b18.py 50 2>&1 | grep "Mev/s" | sed "s/^/[backoff $b] /" | cut -c1-330
done
DFD_POLL_BACKOFF=2 DFD_POOL=1 python tools/gpu_probe.py parity 2>&1 | tail -9 | cut -c1-200
