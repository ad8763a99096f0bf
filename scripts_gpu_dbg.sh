#!/bin/bash
timeout 300 python benchmarks/conv_bench.py --layers s2d_proxy --variants 128:1:1,128:1:2,128:1:8 --batch 8 2>&1 | tail -3
timeout 300 python benchmarks/conv_bench.py --layers 1 --variants 64:1:4 --batch 8 2>&1 | tail -1
