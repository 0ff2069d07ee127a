#!/bin/bash
# pooled event-loop kernel (dfd_run_kernel_pool) vs the static one-warp-per-block assignment; GPU parity suite first
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
DFD_PROFILE_PRINT=1 python tools/gpu_probe.py big131 2>&1 | grep "Mev/s\|cycles per warp" | sed "s/^/[pool] /" | cut -c1-260
DFD_POOL=0 python tools/gpu_probe.py big131 2>&1 | grep "Mev/s" | sed "s/^/[static] /" | cut -c1-260
python tools/gpu_probe.py perf 2>&1 | grep "Mev/s" | sed "s/^/[small] /" | cut -c1-260
python tools/gpu_probe.py big1m 2>&1 | grep "Mev/s" | sed "s/^/[1m pool] /" | cut -c1-260
DFD_POOL=0 python tools/gpu_probe.py big1m 2>&1 | grep "Mev/s" | sed "s/^/[1m static] /" | cut -c1-260
