#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/gpu_probe.py big131 2>&1 | grep "Mev/s" | sed "s/^/[prefetch] /" | cut -c1-260
DFD_PREFETCH=0 python tools/gpu_probe.py big131 2>&1 | grep "Mev/s" | sed "s/^/[no prefetch] /" | cut -c1-260
DFD_PREFETCH=1 python tools/gpu_probe.py perf 2>&1 | grep "Mev/s" | sed "s/^/[small prefetch] /" | cut -c1-260
python tools/gpu_probe.py big1m 2>&1 | grep "Mev/s" | sed "s/^/[1m prefetch] /" | cut -c1-260
