#!/bin/bash
cd "$(dirname "$0")/.."
for g in 592 1184 1776 2368; do
DFD_GRID=$g python tools/gpu_probe.py big131 2>&1 | grep "Mev/s" | sed -n 2p | sed "s/^/[grid $g] /" | cut -c1-230
done
