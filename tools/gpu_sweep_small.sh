#!/bin/bash
# isolated A/B (same box, graph replay) of split-K factors / tile widths for the M <= 1536 layers
mkdir -p gpurun_out
O=gpurun_out/sweep_small.txt
: > $O
run() { timeout 300 python benchmarks/conv_bench.py --batch 8 --iters 20 "$@" >> $O 2>&1; }
run --layers 5,5_1 --variants "256:6:1,256:4:1,256:5:1,256:8:1,256:12:1,128:3:1,128:4:1,128:6:1,256:3:2,256:6:2"
run --layers 6,6_1 --variants "256:8:1,256:6:1,256:12:1,256:9:1,128:4:1,128:6:1,128:8:1,256:4:2,256:8:2"
run --layers deconv5 --variants "64:1:36,128:1:36,64:1:34,128:1:34"
run --layers deconv4 --variants "128:1:36,64:1:36,128:1:34,64:1:34"
run --layers 4_1 --variants "192:1:1,256:1:1,128:1:1,128:1:2,256:2:1"
run --layers 4 --variants "192:1:2,192:1:1,256:1:1,128:1:1,128:1:2"
cat $O | cut -c1-120
