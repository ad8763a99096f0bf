#!/bin/bash
# A/B of per-layer tilings through OFS_TUNE; prints the per-kernel breakdown for each setting
mkdir -p gpurun_out
i=0
while IFS= read -r tune; do
  i=$((i+1))
  echo "=== OFS_TUNE=$tune"
  OFS_TUNE="$tune" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/tune_$i.log 2>&1
  python - "$i" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/tune_{sys.argv[1]}.log").read().strip().splitlines()[-1])
    print("value %.0f pairs/s  ms/step %.4f  conv %.0f TF" % (d["value"], d["ms_per_step"], d["roofline"]["achieved"]))
    print("  " + "  ".join("%s=%.1f" % (b["kernel"].replace("gemm:", ""), b["ms"] * 1e3) for b in d["breakdown"] if b["kernel"].startswith("gemm")))
except Exception as e:
    print("failed", e); print(open(f"gpurun_out/tune_{sys.argv[1]}.log").read()[-1500:])
PY
done < "$1"
