#!/bin/bash
# 2 GPUs (gpurun --gpus 2): multi-process parity check over CUDA-IPC incl. the snapshot records, then the sharded random campaign
cd "$(dirname "$0")/.."
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py 2>&1 | grep "GPUs\]\|MGPU\|differs\|Error\|error" | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/fuzz_parity.py mgpu --cases 120 --seed 3201 2>&1 | grep "MISMATCH\|^mgpu" | cut -c1-600
