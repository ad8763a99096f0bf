#!/bin/bash
# per-item stamps of both single-thread roles (conv_bench debug bit 256) + skip-everything handshake rate (debug 7)
mkdir -p gpurun_out
O=gpurun_out/fine_trace.txt
: > $O
run() { timeout 300 python benchmarks/conv_bench.py --batch 8 --iters 20 "$@" >> $O 2>&1; }
run --layers deconv3 --variants "128:1:34:256" --trace
run --layers deconv2 --variants "64:1:34:256" --trace
run --layers 3_1 --variants "256:1:1:256" --trace
run --layers 4_1 --variants "192:1:1:256" --trace
run --layers deconv3 --variants "128:1:34:0,128:1:34:7,128:1:34:1,128:1:34:6,128:1:8:0"
run --layers deconv2 --variants "64:1:34:0,64:1:34:7,64:1:8:0,64:1:8:7"
run --layers 3_1 --variants "256:1:1:0,256:1:1:7"
cut -c1-400 $O
