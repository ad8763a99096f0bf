#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q -x --timeout 600 --tb=short -k "slab" 2>&1 | tail -6 > gpurun_out/t_slab.log
cat gpurun_out/t_slab.log
O=gpurun_out/conv1x2_bench.txt
: > $O
timeout 300 python benchmarks/conv_bench.py --batch 8 --iters 20 --layers 1 --variants "64:1:4,128:1:5" --trace >> $O 2>&1
cut -c1-330 $O
: > gpurun_out/ab_c1x2.jsonl
timeout 300 python benchmarks/layer_ab.py conv1_one_pixel >> gpurun_out/ab_c1x2.jsonl 2> gpurun_out/ab_c1x2.err
OFS_CONV1X2=1 timeout 300 python benchmarks/layer_ab.py conv1_two_pixel_unpadded >> gpurun_out/ab_c1x2.jsonl 2>> gpurun_out/ab_c1x2.err
timeout 300 python benchmarks/layer_ab.py conv1_one_pixel >> gpurun_out/ab_c1x2.jsonl 2>> gpurun_out/ab_c1x2.err
OFS_CONV1X2=1 timeout 300 python benchmarks/layer_ab.py conv1_two_pixel_unpadded >> gpurun_out/ab_c1x2.jsonl 2>> gpurun_out/ab_c1x2.err
cut -c1-200 gpurun_out/ab_c1x2.jsonl; tail -3 gpurun_out/ab_c1x2.err
OFS_CONV1X2=1 timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -q -x --timeout 600 --tb=short -k "network or batch_independence or full_size" 2>&1 | tail -3
