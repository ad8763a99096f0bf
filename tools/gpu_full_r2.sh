#!/bin/bash
# Full validation pass: every GPU test, the default bench line, the ncu launch list of the same command.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 --tb=short 2>&1 | tail -25 > gpurun_out/t_all.log
tail -12 gpurun_out/t_all.log
timeout 900 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err
tail -c 6000 gpurun_out/bench_default.log
tail -5 gpurun_out/bench_default.err
OFS_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --sustained-seconds 0 > gpurun_out/ncu_list.log 2>&1
wc -l gpurun_out/launches.csv
