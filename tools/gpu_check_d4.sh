#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q -x --timeout 600 --tb=short 2>&1 | tail -12 > gpurun_out/t_net.log
cat gpurun_out/t_net.log
O=gpurun_out/d4_layers.txt
: > $O
run() { timeout 300 python benchmarks/conv_bench.py --batch 8 --iters 20 "$@" >> $O 2>&1; }
run --layers deconv4 --variants "128:1:36,128:1:34,128:2:34,128:3:34,128:4:34,64:2:34,64:3:34"
run --layers deconv5 --variants "64:1:34,128:2:34,128:3:34,64:2:34,128:4:34"
run --layers deconv3 --variants "128:1:36,128:2:34,128:3:34"
cut -c1-100 $O
: > gpurun_out/ab_d4.jsonl
OFS_TUNE=deconv4:128:1:2 timeout 300 python benchmarks/layer_ab.py d4_pairs >> gpurun_out/ab_d4.jsonl 2> gpurun_out/ab_d4.err
timeout 300 python benchmarks/layer_ab.py d4_splitk3 >> gpurun_out/ab_d4.jsonl 2>> gpurun_out/ab_d4.err
cut -c1-420 gpurun_out/ab_d4.jsonl; tail -3 gpurun_out/ab_d4.err
