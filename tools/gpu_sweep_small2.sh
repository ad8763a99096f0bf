#!/bin/bash
# narrower tiles for the M <= 1536 layers: smaller fp32 partial tiles per CTA (the partial epilogue is as long as the K loop)
mkdir -p gpurun_out
O=gpurun_out/sweep_small2.txt
: > $O
run() { timeout 300 python benchmarks/conv_bench.py --batch 8 --iters 20 "$@" >> $O 2>&1; }
run --layers 5,5_1 --variants "256:6:1,64:1:1,64:1:2,64:2:1,64:2:2,128:1:1,128:1:2,128:2:1,128:2:2,128:3:1"
run --layers 6 --variants "128:4:1,64:1:1,64:2:1,64:3:1,64:4:1,128:2:1,128:3:1,64:2:2"
run --layers 6_1 --variants "128:4:1,64:2:1,64:3:1,64:4:1,64:6:1,128:3:1,128:6:1,64:3:2"
cat $O | cut -c1-120
