#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -q --timeout 200 --tb=short -k "conv_gemm" 2>&1 | tail -15
bash scripts_gpu_tune.sh tunes.txt
