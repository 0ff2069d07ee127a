#!/bin/bash
# Round-end profile capture (run under gpurun on ONE GPU).  Every program is run plainly first and must exit 0
# before it is run under ncu; results land in gpurun_out/.  Read them here with tools/profile_extract.py.
set -u
OUT=gpurun_out/profile
mkdir -p $OUT
python tools/gen_topology.py 128 4 /tmp/huge --name huge > /dev/null
# 1. plain runs
python bench.py --steps 2 --warmup 1 > $OUT/bench_plain.json 2> $OUT/bench_plain.err || { echo "bench failed"; tail -5 $OUT/bench_plain.err; exit 1; }
python tools/phase_probe.py configs/dfdally-1056.conf num_messages=100 > $OUT/probe_1056.txt || exit 1
python tools/phase_probe.py configs/dfdally-8320-par.conf num_messages=50 > $OUT/probe_8320.txt || exit 1
python tools/phase_probe.py /tmp/huge/huge.conf num_messages=10 > $OUT/probe_1m.txt || exit 1
# 2. launch list of the bench command
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 > $OUT/ncu_launches.log 2>&1
# 3. full capture of the window kernel on the bench workload (second launch: the first one is phase_probe's warm-up run)
ncu --set full --clock-control none --import-source on -k regex:dfd_run_kernel -s 1 -c 1 -f -o $OUT/window_kernel_1056 python tools/phase_probe.py configs/dfdally-1056.conf num_messages=100 > $OUT/ncu_full.log 2>&1
# 4. DRAM traffic of the window kernel at scale (8,320- and 1,050,624-terminal configs)
M=dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none -k regex:dfd_run_kernel -s 1 -c 1 --csv --log-file $OUT/dram_8320.csv python tools/phase_probe.py configs/dfdally-8320-par.conf num_messages=50 > $OUT/ncu_8320.log 2>&1
ncu --metrics $M --clock-control none -k regex:dfd_run_kernel -s 1 -c 1 --csv --log-file $OUT/dram_1m.csv python tools/phase_probe.py /tmp/huge/huge.conf num_messages=10 > $OUT/ncu_1m.log 2>&1
ls -la $OUT
