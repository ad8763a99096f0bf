#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/probe_fine2.log
: > $O
run() { timeout 300 python benchmarks/conv_bench.py --batch 8 --iters 10 --trace "$@" >> $O 2>&1; }
run --layers deconv3 --variants "128:1:1:256,128:1:1:257,128:1:1:262,128:1:2:256"
run --layers deconv2 --variants "64:1:1:256,64:1:1:257,64:1:1:262"
run --layers 3 --variants "256:1:2:256,256:1:2:257,256:1:2:262"
run --layers 4_1 --variants "192:1:1:256"
run --layers 5_1 --variants "256:6:1:256"
run --layers 6_1 --variants "256:8:1:256"
run --layers 1 --variants "64:1:4:256"
run --layers 2 --variants "128:1:4:256"
grep -v "^           \(previous\|SM clock\|epilogue\|per stage\)" $O | cut -c1-460
