#!/bin/bash
# round-2 measurement batch: probes (no profiler) first, then ncu captures of the same commands
set -x
cd "$(dirname "$0")/.."
python tools/gpu_probe.py perf > gpurun_out/probe5.log 2>&1
python tools/gpu_probe.py big > gpurun_out/probe_big2.log 2>&1
python tools/gen_topology.py 64 2 /tmp/big131 --name big --routing prog-adaptive > /dev/null
python tools/run_one.py configs/dfdally-8320-par.conf num_messages=6 > gpurun_out/plain_r2c.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dfd_run_kernel -o gpurun_out/r2c_8320 -f python tools/run_one.py configs/dfdally-8320-par.conf num_messages=6 > gpurun_out/ncu_r2c.log 2>&1
python tools/run_one.py /tmp/big131/big.conf num_messages=2 > gpurun_out/plain_r2d.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dfd_run_kernel -o gpurun_out/r2d_131k -f python tools/run_one.py /tmp/big131/big.conf num_messages=2 > gpurun_out/ncu_r2d.log 2>&1
tail -3 gpurun_out/probe5.log gpurun_out/probe_big2.log gpurun_out/plain_r2c.log gpurun_out/plain_r2d.log
