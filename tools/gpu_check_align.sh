#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 --tb=short 2>&1 | tail -5 > gpurun_out/t_all.log
cat gpurun_out/t_all.log
timeout 300 python benchmarks/conv_bench.py --batch 8 --iters 20 --layers predict2,predict2t --variants "32:1:32" 2>&1 | cut -c1-100
: > gpurun_out/ab_align.jsonl
timeout 300 python benchmarks/layer_ab.py aligned_concat_strides >> gpurun_out/ab_align.jsonl 2> gpurun_out/ab_align.err
timeout 300 python benchmarks/layer_ab.py aligned_concat_strides >> gpurun_out/ab_align.jsonl 2>> gpurun_out/ab_align.err
cut -c1-700 gpurun_out/ab_align.jsonl; tail -3 gpurun_out/ab_align.err
