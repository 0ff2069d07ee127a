#!/bin/bash
cd "$(dirname "$0")/.."
python tools/gpu_probe.py perf big131 2>&1 | grep "Mev/s" | sed "s/^/[sleep2000] /" | cut -c1-250
DFD_IDLE_SLEEP_NS=0 python tools/gpu_probe.py big131 2>&1 | grep "Mev/s" | sed "s/^/[sleep0] /" | cut -c1-250
DFD_IDLE_SLEEP_NS=8000 python tools/gpu_probe.py big131 2>&1 | grep "Mev/s" | sed "s/^/[sleep8000] /" | cut -c1-250
