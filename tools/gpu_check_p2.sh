#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q -x --timeout 600 --tb=short -k "not tilings" 2>&1 | tail -15 > gpurun_out/t_net.log
tail -8 gpurun_out/t_net.log
timeout 600 python -m pytest tests/test_gpu_clip.py tests/test_gpu_modes.py -m gpu -q -x --timeout 600 --tb=short 2>&1 | tail -5
O=gpurun_out/ab_p2.jsonl
: > $O
run() { timeout 300 python benchmarks/layer_ab.py "$1" >> $O 2>> gpurun_out/ab_p2.err; }
run p2_fused
OFS_P2_FUSED=0 run p2_gemm_gather
run p2_fused_again
cat $O
tail -5 gpurun_out/ab_p2.err
