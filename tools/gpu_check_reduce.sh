#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q -x --timeout 600 --tb=short 2>&1 | tail -4
timeout 300 python benchmarks/conv_bench.py --batch 8 --iters 20 --layers 5,5_1,6,6_1 2>&1 | cut -c1-100
timeout 300 python benchmarks/layer_ab.py reduce_batched_loads 2> gpurun_out/ab_red.err | cut -c1-330
