#!/bin/bash
timeout 300 python benchmarks/conv_bench.py --layers s2d_proxy --variants 128:1:1:0,128:1:1:1,128:1:1:6,128:1:1:8,128:1:1:2,128:1:1:4 --batch 8 2>&1 | tail -6
timeout 300 python benchmarks/conv_bench.py --layers s2d_proxy --variants 128:1:1 --trace --batch 8 2>&1 | tail -6
