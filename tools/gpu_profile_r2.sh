#!/bin/bash
# round-2 evidence run: all GPU tests, smoke, the default bench line, then the ncu passes of the same command
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 --tb=short 2>&1 | tail -15 > gpurun_out/t_all.log
tail -6 gpurun_out/t_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err
tail -c 3000 gpurun_out/bench_default.log
timeout 600 python bench.py --streams 1 --no-cpu-baseline --sustained-seconds 0 > gpurun_out/bench_streams1.log 2>&1
CMD="env OFS_GRAPH=0 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --sustained-seconds 0"
$CMD > gpurun_out/plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm|deconv_stack|splitk_reduce" -s 38 -c 19 -o gpurun_out/prof_conv -f $CMD > gpurun_out/ncu_conv.log 2>&1
$CMD > gpurun_out/plain3.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"warp5|pack_act|pack27|predict2_gather|pyr_kernel" -s 14 -c 7 -o gpurun_out/prof_misc -f $CMD > gpurun_out/ncu_misc.log 2>&1
tail -2 gpurun_out/ncu_launches.log gpurun_out/ncu_conv.log gpurun_out/ncu_misc.log
ls -la gpurun_out | head -40
