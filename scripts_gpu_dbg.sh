#!/bin/bash
mkdir -p gpurun_out
timeout 300 python benchmarks/conv_bench.py --layers 5_1 --variants 256:5:16,256:4:16 --trace --batch 8 2>&1 | tail -14
timeout 300 python benchmarks/conv_bench.py --layers 6_1 --variants 256:6:16 --trace --batch 8 2>&1 | tail -7
