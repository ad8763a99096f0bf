#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_clip.py -m gpu -q --timeout 200 --tb=short 2>&1 | tail -3
timeout 200 python benchmarks/clip_bench.py 2>&1 | grep "720x1280" | tail -5
