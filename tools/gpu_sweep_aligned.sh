#!/bin/bash
# tiling choices re-checked on the 64-channel-aligned concat strides
mkdir -p gpurun_out
O=gpurun_out/sweep_aligned.txt
: > $O
run() { timeout 300 python benchmarks/conv_bench.py --batch 8 --iters 20 "$@" >> $O 2>&1; }
run --layers 2 --variants "128:1:4"
run --layers 3 --variants "256:1:2,256:1:1,128:1:2"
run --layers 3_1 --variants "256:1:1,256:1:2,128:1:1"
run --layers 4 --variants "192:1:2,192:1:1,256:1:2,256:1:1,128:1:2"
run --layers 4_1 --variants "192:1:1,192:1:2,256:1:1,256:1:2"
run --layers 5,5_1 --variants "256:6:1,256:4:1,256:5:1,128:3:1,256:3:2"
run --layers 6,6_1 --variants "128:4:1,256:8:1,128:6:1,128:3:1,256:6:1"
run --layers deconv5 --variants "64:1:34,64:1:36,128:1:34,128:1:36"
run --layers deconv4 --variants "128:1:36,128:1:34,128:3:34,64:1:36,128:2:34"
run --layers deconv3 --variants "128:1:36,128:1:34,64:1:36,64:1:34"
run --layers deconv2 --variants "64:1:66,64:1:64,64:1:36,64:1:34"
run --layers predict2 --variants "32:1:32,32:1:1,16:1:1"
cut -c1-100 $O
