#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -q --timeout 300 --tb=short -k "tilings" 2>&1 | tail -15 > gpurun_out/t_tilings.log
tail -8 gpurun_out/t_tilings.log
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q -x --timeout 600 --tb=short -k "not tilings" 2>&1 | tail -15 > gpurun_out/t_net.log
tail -8 gpurun_out/t_net.log
O=gpurun_out/conv1_bench.txt
: > $O
timeout 300 python benchmarks/conv_bench.py --batch 8 --iters 20 --layers 1 --variants "64:1:4,128:1:5" --trace >> $O 2>&1
cat $O | cut -c1-300
O=gpurun_out/ab_conv1.jsonl
: > $O
run() { timeout 300 python benchmarks/layer_ab.py "$1" >> $O 2>> gpurun_out/ab_conv1.err; }
run conv1x2
OFS_CONV1X2=0 run conv1_one_pixel
run conv1x2_again
cat $O
tail -5 gpurun_out/ab_conv1.err
