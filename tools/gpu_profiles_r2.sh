#!/bin/bash
# round-2 profile artefacts for profiles/: plain bench first, then the ncu launch list of the same command, then one
# --set full capture of the event-loop kernel on the bench workload itself (131,584 terminals, PAR, uniform, 10 messages:
# the pooled kernel) and one on BASELINE configs[2] (8,320 terminals, PAR: one warp per group), each after a plain run of
# the same command has exited 0
cd "$(dirname "$0")/.."
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_plain_bench.json 2> gpurun_out/r2_plain_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_ncu_bench.log 2>&1
python tools/gen_topology.py 64 2 /tmp/big131 --name big --routing prog-adaptive > /dev/null
python tools/run_one.py /tmp/big131/big.conf num_messages=10 > gpurun_out/r2_plain_one.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dfd_run_kernel -o gpurun_out/r2_eventloop_131k -f python tools/run_one.py /tmp/big131/big.conf num_messages=10 > gpurun_out/r2_ncu_full.log 2>&1
python tools/run_one.py configs/dfdally-8320-par.conf num_messages=20 > gpurun_out/r2_plain_8320.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dfd_run_kernel -o gpurun_out/r2_eventloop_8320 -f python tools/run_one.py configs/dfdally-8320-par.conf num_messages=20 > gpurun_out/r2_ncu_8320.log 2>&1
cut -c1-400 gpurun_out/r2_plain_bench.json; cat gpurun_out/r2_plain_one.log gpurun_out/r2_plain_8320.log; tail -n 3 gpurun_out/r2_launches.csv | cut -c1-300
