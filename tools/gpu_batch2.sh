#!/bin/bash
cd "$(dirname "$0")/.."
DFD_PROFILE_PRINT=1 python tools/gpu_probe.py perf big131 2>&1 | grep "Mev/s\|by cause\|rounds" | sed "s/^/[base] /"
DFD_ACT_BUDGET=16 python tools/gpu_probe.py perf big131 2>&1 | grep "Mev/s" | sed 's/^/[budget16] /'
DFD_ACT_BUDGET=64 python tools/gpu_probe.py big131 2>&1 | grep "Mev/s" | sed 's/^/[budget64] /'
DFD_MINBLOCKS=32 python tools/gpu_probe.py big131 2>&1 | grep "Mev/s" | sed 's/^/[mb32] /'
DFD_MINBLOCKS=32 DFD_ACT_BUDGET=16 python tools/gpu_probe.py big131 2>&1 | grep "Mev/s" | sed 's/^/[mb32 budget16] /'
