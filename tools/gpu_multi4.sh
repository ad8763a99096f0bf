#!/bin/bash
# 4-GPU box: the NCCL gather test, then BASELINE configs[3] at N = 2 and 4
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --timeout 500 --tb=short 2>&1 | tail -8 > gpurun_out/t_multi.log
cat gpurun_out/t_multi.log
for n in 2 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
    benchmarks/config4_clips.py --json gpurun_out/config4_n$n.json > gpurun_out/config4_n$n.log 2>&1
  tail -2 gpurun_out/config4_n$n.log | cut -c1-1500
done
