#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_modes.py -m gpu -q -x --timeout 600 --tb=short 2>&1 | tail -6 > gpurun_out/t_net.log
cat gpurun_out/t_net.log
: > gpurun_out/ab_p2g.jsonl
timeout 300 python benchmarks/layer_ab.py p2gather_branchfree >> gpurun_out/ab_p2g.jsonl 2> gpurun_out/ab_p2g.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/ab_p2g.jsonl").read().strip().splitlines()[-1])
print(d["pairs_s_1"], d["pairs_s_2"], d["dense_ms"], {k: v for k, v in d["us"].items() if not k[0].isdigit() and not k.startswith("deconv")})
PY
