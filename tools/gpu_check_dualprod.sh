#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q -x --timeout 600 --tb=short 2>&1 | tail -8 > gpurun_out/t_net.log
tail -5 gpurun_out/t_net.log
O=gpurun_out/dualprod_layers.txt
: > $O
run() { timeout 300 python benchmarks/conv_bench.py --batch 8 --iters 20 "$@" >> $O 2>&1; }
run --layers 1 --variants "64:1:4"
run --layers 2 --variants "128:1:4"
run --layers 3 --variants "256:1:2,256:1:1"
run --layers 3_1 --variants "256:1:1,256:1:2"
run --layers 4 --variants "192:1:2,192:1:1"
run --layers 4_1 --variants "192:1:1,192:1:2"
run --layers 5,5_1 --variants "256:6:1"
run --layers 6,6_1 --variants "128:4:1,256:8:1"
run --layers deconv5 --variants "64:1:34,64:1:36"
run --layers deconv4 --variants "128:1:36,128:1:34"
run --layers deconv3 --variants "128:1:36,128:1:34"
run --layers deconv2 --variants "64:1:66,64:1:64,64:1:34"
run --layers predict2 --variants "32:1:1"
run --layers deconv3 --variants "128:1:34" --trace
run --layers 3_1 --variants "256:1:1" --trace
cut -c1-330 $O
timeout 300 python benchmarks/layer_ab.py dual_producers > gpurun_out/ab_dualprod.jsonl 2> gpurun_out/ab_dualprod.err; cat gpurun_out/ab_dualprod.jsonl; tail -3 gpurun_out/ab_dualprod.err
