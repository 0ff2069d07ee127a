#!/bin/bash
# seeded random parity campaign on the build with R_SNAPSHOT events (about every fifth case draws router_buffer_snapshots)
cd "$(dirname "$0")/.."
timeout 900 python tools/fuzz_parity.py gpu --cases 300 --seed 3101 2>&1 | grep "MISMATCH\|^gpu:" | cut -c1-600
DFD_POOL=1 timeout 900 python tools/fuzz_parity.py gpu --cases 300 --seed 3102 2>&1 | grep "MISMATCH\|^gpu:" | cut -c1-600
DFD_WIDE_REGS=0 timeout 900 python tools/fuzz_parity.py gpu --cases 150 --seed 3103 2>&1 | grep "MISMATCH\|^gpu:" | cut -c1-600
