#!/bin/bash
# ncu --set full captures of the event-loop kernel with the current build: 8,320 terminals PAR (one warp per group, latency-bound)
# and 131,584 terminals PAR (pooled kernel); each after a plain run of the same command
cd "$(dirname "$0")/.."
python tools/run_one.py configs/dfdally-8320-par.conf num_messages=20 > gpurun_out/r2g_plain_8320.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dfd_run_kernel -o gpurun_out/r2g_8320 -f python tools/run_one.py configs/dfdally-8320-par.conf num_messages=20 > gpurun_out/r2g_ncu_8320.log 2>&1
python tools/gen_topology.py 64 2 /tmp/big131 --name big --routing prog-adaptive > /dev/null
python tools/run_one.py /tmp/big131/big.conf num_messages=2 > gpurun_out/r2g_plain_131k.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dfd_run_kernel -o gpurun_out/r2g_131k -f python tools/run_one.py /tmp/big131/big.conf num_messages=2 > gpurun_out/r2g_ncu_131k.log 2>&1
cat gpurun_out/r2g_plain_8320.log gpurun_out/r2g_plain_131k.log
