#!/bin/bash
# where does a layer's time go?  per-CTA role timeline + stall counters (conv_bench --trace), and the debug skip switches
mkdir -p gpurun_out
O=gpurun_out/trace_layers.txt
: > $O
run() { timeout 300 python benchmarks/conv_bench.py --batch 8 --iters 20 "$@" >> $O 2>&1; }
run --layers 3_1 --variants "256:1:1,256:1:2" --trace
run --layers 3_1 --variants "256:1:1:1,256:1:1:6,256:1:1:8"
run --layers 4_1 --variants "192:1:1,192:1:2" --trace
run --layers 4 --variants "192:1:1,192:1:2" --trace
run --layers 3 --variants "256:1:1,256:1:2" --trace
run --layers deconv2 --variants "64:1:34,64:1:36" --trace
run --layers deconv2 --variants "64:1:34:1,64:1:34:6,64:1:34:8,64:1:64,64:1:66"
run --layers deconv3 --variants "128:1:34,128:1:36" --trace
run --layers deconv4 --variants "128:1:34,128:1:36" --trace
run --layers 5_1 --variants "256:6:1" --trace
run --layers 6_1 --variants "256:8:1" --trace
run --layers 1 --variants "64:1:4" --trace
run --layers 2 --variants "128:1:4" --trace
cat $O | cut -c1-260
