#!/bin/bash
# publisher queues with 128-word batches: parity of the sharded tests, then direct fence against publishers with the ranks on one GPU
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sharded" 2>&1 | tail -2
for n in 0 2 8; do
DFD_NPUB=$n timeout 120 python tools/sharded_local_probe.py 2 64 16 2>&1 | tail -1 | cut -c1-220
done
