#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -q --timeout 200 --tb=short -k "conv or slab" 2>&1 | tail -8
timeout 300 python benchmarks/conv_bench.py --layers 1 --variants 64:1:4:0 --batch 8 2>&1 | tail -4
timeout 300 python benchmarks/conv_bench.py --layers 2 --variants 128:1:4:0 --batch 8 2>&1 | tail -4
timeout 300 python benchmarks/conv_bench.py --layers 3 --variants 256:1:2:0 --batch 8 2>&1 | tail -4
timeout 300 python bench.py --steps 100 --warmup 5 2>&1 | tail -3
