#!/bin/bash
# 8-GPU box: BASELINE configs[3] at N = 8, 4, 2 and 1 (strong scaling of 8 x 1024 frames), NCCL gather of the uint8 output
mkdir -p gpurun_out
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
    benchmarks/config4_clips.py --json gpurun_out/config4_n$n.json > gpurun_out/config4_n$n.log 2>&1
  tail -1 gpurun_out/config4_n$n.log | cut -c1-1200
done
timeout 600 python benchmarks/config4_clips.py --json gpurun_out/config4_n1.json > gpurun_out/config4_n1.log 2>&1
tail -1 gpurun_out/config4_n1.log | cut -c1-1200
