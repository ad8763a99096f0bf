#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -q -x --timeout 300 --tb=short -k "chain" 2>&1 | tail -15 > gpurun_out/t_chain.log
cat gpurun_out/t_chain.log
: > gpurun_out/ab_chain.jsonl
OFS_CHAIN=0 timeout 300 python benchmarks/layer_ab.py chain_off >> gpurun_out/ab_chain.jsonl 2> gpurun_out/ab_chain.err
OFS_CHAIN=1 timeout 300 python benchmarks/layer_ab.py chain_on >> gpurun_out/ab_chain.jsonl 2>> gpurun_out/ab_chain.err
OFS_CHAIN=0 timeout 300 python benchmarks/layer_ab.py chain_off >> gpurun_out/ab_chain.jsonl 2>> gpurun_out/ab_chain.err
OFS_CHAIN=1 timeout 300 python benchmarks/layer_ab.py chain_on >> gpurun_out/ab_chain.jsonl 2>> gpurun_out/ab_chain.err
cat gpurun_out/ab_chain.jsonl; tail -5 gpurun_out/ab_chain.err
