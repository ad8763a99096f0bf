#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_clip.py -m gpu -q --timeout 300 --tb=short 2>&1 | tail -25
timeout 300 python benchmarks/clip_bench.py 2>&1 | tail -12
timeout 300 python bench.py --steps 100 --warmup 5 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e_clip_driver'])"
