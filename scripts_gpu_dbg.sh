#!/bin/bash
mkdir -p gpurun_out
for t in 0 1 0 1; do
OFS_TAIL_HALF=$t timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('tail', $t, round(d['value']), round(d['value_one_step_at_a_time']), round(d['roofline']['frac'],4), [ (b['kernel'],round(b['ms']*1e3,1)) for b in d['breakdown'] if b['kernel'] in ('gemm:3','gemm:3_1')])"
done
timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -q --timeout 300 --tb=short -k "not conv_gemm" 2>&1 | tail -4
