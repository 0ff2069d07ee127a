#!/bin/bash
cd "$(dirname "$0")/.."
run() { python tools/gpu_probe.py perf big131 2>&1 | grep "Mev/s" | sed -n '3,4p;6,7p' | sed "s/^/[$1] /" | cut -c1-150; }
DFD_HYST_FRAC=0.1 run hyst0.1
DFD_HYST_FRAC=0.5 run hyst0.5
DFD_HYST_FRAC=0.8 run hyst0.8
DFD_ACT_BUDGET=8 run budget8
DFD_ACT_BUDGET=2 run budget2
DFD_ACT_BUDGET=32 run budget32
