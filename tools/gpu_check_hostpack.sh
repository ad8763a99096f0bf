#!/bin/bash
mkdir -p gpurun_out
nproc; lscpu | grep -E "Model name|^CPU\(s\)|Flags" | cut -c1-200 | head -3
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q -x --timeout 600 --tb=short -k "host or stabilize or full_size or precision" -s 2>&1 | tail -25 > gpurun_out/t_host.log
tail -25 gpurun_out/t_host.log
for v in "default" "OFS_HOST_PACK=0" "OFS_HOST_PACK_THREADS=3" "OFS_HOST_PACK_THREADS=8"; do
  if [ "$v" == "default" ]; then e=""; else e="$v"; fi
  env $e timeout 600 python bench.py --steps 50 --no-cpu-baseline --sustained-seconds 0 > gpurun_out/bench_hp.log 2>&1
  python - "$v" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/bench_hp.log").read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"], 1), "h2d", d["e2e"]["h2d_bytes_per_step"], "clip", round(d["e2e_clip_driver"]["value"]))
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_hp.log").read()[-2000:])
PY
done
