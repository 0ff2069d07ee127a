#!/bin/bash
cd "$(dirname "$0")/.."
DFD_PROFILE_PRINT=1 python tools/gpu_probe.py perf big131 2>&1 | grep "Mev/s\|by cause\|rounds" | cut -c1-330
