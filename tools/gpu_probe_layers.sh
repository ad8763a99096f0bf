#!/bin/bash
# Per-layer conv_bench of the network's own tilings + "skip MMA" (dbg 1) / "skip loads" (dbg 6) variants with traces:
# which resource (tensor pipe, L2 -> shared-memory feed, wave quantisation) binds each layer today.
mkdir -p gpurun_out
O=gpurun_out/probe_layers.log
: > $O
run() { timeout 300 python benchmarks/conv_bench.py --batch 8 --iters 20 --trace "$@" >> $O 2>&1; }
run --layers 1 --variants "64:1:4,64:1:4:1,64:1:4:6"
run --layers 2 --variants "128:1:4,128:1:4:1,128:1:4:6"
run --layers 3 --variants "256:1:2,256:1:2:1,256:1:2:6,256:1:1,256:1:1:1,256:1:1:6"
run --layers 3_1 --variants "256:1:1,256:1:1:1,256:1:1:6,256:1:2,256:1:2:1,256:1:2:6"
run --layers 4 --variants "192:1:1,192:1:1:1,192:1:1:6,256:1:2,256:1:1"
run --layers 4_1 --variants "192:1:1,192:1:1:1,192:1:1:6,256:1:2,256:1:2:1,256:1:2:6,256:1:1"
run --layers 5,5_1 --variants "256:6:1,256:6:1:1,256:6:1:6,256:6:2"
run --layers 6,6_1 --variants "256:8:1,256:8:1:1,256:8:1:6,256:8:2"
run --layers deconv5 --variants "64:1:1,128:1:1,128:1:1:1,128:1:1:6"
run --layers deconv4 --variants "128:1:1,128:1:1:1,128:1:1:6"
run --layers deconv3 --variants "128:1:1,128:1:1:1,128:1:1:6,128:1:2"
run --layers deconv2 --variants "64:1:1,64:1:1:1,64:1:1:6,64:1:2"
grep -v "^  " $O | cut -c1-120
