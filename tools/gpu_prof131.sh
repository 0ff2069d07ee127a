#!/bin/bash
# ncu capture of the bench workload (131,584 terminals, PAR, uniform) with fewer messages; the plain run comes first
cd "$(dirname "$0")/.."
TAG=${1:-r2e}
python tools/gen_topology.py 64 2 /tmp/big131 --name big --routing prog-adaptive > /dev/null
python tools/run_one.py /tmp/big131/big.conf num_messages=2 > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dfd_run_kernel -o gpurun_out/${TAG}_131k -f python tools/run_one.py /tmp/big131/big.conf num_messages=2 > gpurun_out/ncu_$TAG.log 2>&1
cat gpurun_out/plain_$TAG.log
