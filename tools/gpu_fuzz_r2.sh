#!/bin/bash
# seeded random parity campaign on the final round-2 build: one warp per group, then the pooled kernel forced on other cases
cd "$(dirname "$0")/.."
python tools/fuzz_parity.py gpu --cases 300 --seed 2026 2>&1 | grep "MISMATCH\|^gpu:" | cut -c1-600
python tools/fuzz_parity.py gpu --cases 300 --seed 7 2>&1 | grep "MISMATCH\|^gpu:" | cut -c1-600
DFD_POOL=1 python tools/fuzz_parity.py gpu --cases 300 --seed 4052 2>&1 | grep "MISMATCH\|^gpu:" | cut -c1-600
DFD_POOL=1 python tools/fuzz_parity.py gpu --cases 300 --seed 2026 2>&1 | grep "MISMATCH\|^gpu:" | cut -c1-600
