#!/bin/bash
# A/B of launch modes: plain stream launches, PDL only, graph only, graph + PDL
mkdir -p gpurun_out
for mode in "0 0" "1 0" "0 1" "1 1"; do
  set -- $mode
  echo "=== OFS_PDL=$1 OFS_GRAPH=$2"
  OFS_PDL=$1 OFS_GRAPH=$2 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/ab_$1$2.log 2>&1
  python - "$1$2" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/ab_{sys.argv[1]}.log").read().strip().splitlines()[-1])
    print("value %.0f pairs/s  ms/step %.4f  e2e %.0f launches/step %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["launches_per_step"]))
except Exception as e:
    print("failed", e); print(open(f"gpurun_out/ab_{sys.argv[1]}.log").read()[-2000:])
PY
done
