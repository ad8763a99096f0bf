#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/sweep_kgroup.txt
: > $O
run() { timeout 300 python benchmarks/conv_bench.py --batch 8 --iters 20 "$@" >> $O 2>&1; }
run --layers deconv2 --variants "64:1:34,64:1:40"
run --layers deconv3 --variants "128:1:36,128:1:40,64:1:40"
run --layers deconv4 --variants "128:1:34,128:1:40"
run --layers deconv5 --variants "64:1:34,64:1:40,128:1:40"
run --layers 3_1 --variants "256:1:2,128:1:8"
run --layers 2 --variants "128:1:4,128:1:8,128:1:2"
cut -c1-100 $O
