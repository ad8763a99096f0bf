#!/bin/bash
# 128-byte aligned concat strides vs the current odd strides, layer by layer (current default tilings)
mkdir -p gpurun_out
O=gpurun_out/probe_align.log
: > $O
run() { timeout 300 python benchmarks/conv_bench.py --batch 8 --iters 20 "$@" >> $O 2>&1; }
run --layers 2,2a --variants "128:1:4"
run --layers 3,3a --variants "256:1:2"
run --layers 3_1,3_1a --variants "256:1:1"
run --layers 4,4a --variants "192:1:2"
run --layers 4_1,4_1a --variants "192:1:1"
run --layers 5,5a,5_1,5_1a --variants "256:6:1"
run --layers 6,6a --variants "128:4:1"
run --layers deconv5,deconv5a --variants "64:1:34"
run --layers deconv4,deconv4a,deconv3,deconv3a --variants "128:1:36"
run --layers deconv2,deconv2a --variants "64:1:66"
cat $O | cut -c1-100
