#!/bin/bash
# publisher queues (sharded runs): the ranks-on-one-GPU tests exercise them on a single GPU; then the whole GPU suite and a 1-GPU probe
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sharded" 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/gpu_probe.py perf 2>&1 | grep "Mev/s" | cut -c1-130
