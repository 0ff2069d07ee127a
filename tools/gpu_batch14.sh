#!/bin/bash
cd "$(dirname "$0")/.."
for s in 0 200 500 1000 2000 4000; do
DFD_IDLE_SLEEP_NS=$s python tools/gpu_probe.py perf 2>&1 | grep "Mev/s" | sed "s/^/[sleep $s] /" | cut -c1-200
done
