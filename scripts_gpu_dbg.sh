#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_clip.py -m gpu -q --timeout 300 --tb=short 2>&1 | tail -25
timeout 300 python benchmarks/clip_bench.py 2>&1 | tail -12
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/clip_launches.csv python benchmarks/clip_one.py 8 4 > gpurun_out/clip_ncu.log 2>&1
grep -E "resize_u8|assemble|warp5_u8" gpurun_out/clip_launches.csv | tail -4 | cut -c1-60,200-
