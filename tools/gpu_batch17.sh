#!/bin/bash
cd "$(dirname "$0")/.."
python tools/gpu_probe.py parity 2>&1 | grep -c IDENTICAL
DFD_POOL=1 python tools/gpu_probe.py parity 2>&1 | grep -c IDENTICAL
python tools/gpu_probe.py big131 2>&1 | grep "Mev/s" | cut -c1-250
