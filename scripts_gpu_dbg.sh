#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -q --timeout 300 --tb=short -k "not conv_gemm or tail_half" 2>&1 | tail -4
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['value_one_step_at_a_time']), [ (b['kernel'],round(b['ms']*1e3,1)) for b in d['breakdown'] if 'pyramid' in b['kernel']])"
