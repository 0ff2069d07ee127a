#!/bin/bash
# 2 GPUs: publisher queues over real NVLink -- parity check, sharded random campaign, bench at N=2 (pooled kernel + publisher block)
cd "$(dirname "$0")/.."
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py 2>&1 | grep "GPUs\]\|MGPU\|differs\|Error\|error" | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/fuzz_parity.py mgpu --cases 80 --seed 3301 2>&1 | grep "MISMATCH\|^mgpu" | cut -c1-600
NS=2 bash tools/gpu_scale.sh
DFD_NPUB=0 NS=2 bash tools/gpu_scale.sh | sed 's/^/[DFD_NPUB=0] /'
