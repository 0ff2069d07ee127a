#!/bin/bash
cd "$(dirname "$0")/.."
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py 2>&1 | grep -v "^W\|^\*" | tail -15
python bench.py --steps 3 --warmup 3 2>gpurun_out/bench3_n1.err | tee gpurun_out/bench3_n1.json | cut -c1-1500
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 2>gpurun_out/bench3_n2.err | tee gpurun_out/bench3_n2.json | cut -c1-1500
tail -5 gpurun_out/bench3_n1.err gpurun_out/bench3_n2.err
