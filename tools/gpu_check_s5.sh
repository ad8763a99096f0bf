#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_samplers.py tests/test_gpu_modes.py -m gpu -q -x --timeout 600 --tb=short 2>&1 | tail -8 > gpurun_out/t_s5.log
cat gpurun_out/t_s5.log
timeout 600 python benchmarks/sampler_sweep.py --quick 2>&1 | grep -v "staged\|direct" | grep "720x1280\|2160x3840" | grep "| 8 |\|Affine\|Projective\|transformImage" > gpurun_out/sweep_quick.txt; cat gpurun_out/sweep_quick.txt
