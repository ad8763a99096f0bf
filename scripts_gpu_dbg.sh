#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -q --timeout 120 --tb=short -k "conv or net or predict" 2>&1 | tail -5
timeout 300 python benchmarks/conv_bench.py --layers predict2 --variants 32:1:1:0,32:1:1:15 --batch 8 2>&1 | tail -3
timeout 300 python bench.py --steps 100 --warmup 5 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['value_one_step_at_a_time'], d['roofline']['frac'], [ (b['kernel'],b['ms']) for b in d['breakdown'] if 'predict2' in b['kernel']])"
