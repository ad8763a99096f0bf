#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q -x --timeout 600 --tb=short 2>&1 | tail -6 > gpurun_out/t_net.log
cat gpurun_out/t_net.log
: > gpurun_out/ab_pack.jsonl
OFS_PACK27=0 timeout 300 python benchmarks/layer_ab.py pack_generic >> gpurun_out/ab_pack.jsonl 2> gpurun_out/ab_pack.err
timeout 300 python benchmarks/layer_ab.py pack27 >> gpurun_out/ab_pack.jsonl 2>> gpurun_out/ab_pack.err
cut -c1-220 gpurun_out/ab_pack.jsonl; tail -3 gpurun_out/ab_pack.err
timeout 300 python benchmarks/warp_tune.py 2>&1 | tail -7
