#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -q --timeout 120 --tb=short -k "kgroup or fp32_out" 2>&1 | tail -15
timeout 300 python benchmarks/conv_bench.py --layers predict2 --variants 32:1:1,32:1:32 --batch 8 2>&1 | tail -2
