#!/bin/bash
for lim in 148 74 100 74 148; do
OFS_SM_LIMIT=$lim timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('limit', $lim, round(d['value']), round(d['value_one_step_at_a_time']))"
done
