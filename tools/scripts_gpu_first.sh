#!/bin/bash
# first GPU bring-up: every stage in its own process (a faulting kernel poisons only its own context)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_samplers.py -m gpu -q --timeout 300 --tb=short 2>&1 | tail -40 > gpurun_out/t_samplers.log
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q --timeout 300 --tb=short -k conv_gemm 2>&1 | tail -80 > gpurun_out/t_conv.log
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q --timeout 600 --tb=short -s -k "not conv_gemm" 2>&1 | tail -100 > gpurun_out/t_net.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1
for f in smi.txt t_samplers.log t_conv.log t_net.log smoke.log bench.log; do echo "=== $f"; tail -45 gpurun_out/$f; done
