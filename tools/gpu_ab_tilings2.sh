#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/ab_tilings2.jsonl
: > $O
old() { OFS_STACK=1 OFS_TUNE="3_1:256:1:1,4:192:1:2,deconv4:128:1:2" timeout 300 python benchmarks/layer_ab.py old_tilings >> $O 2>> gpurun_out/ab_tilings2.err; }
new() { timeout 300 python benchmarks/layer_ab.py new_tilings >> $O 2>> gpurun_out/ab_tilings2.err; }
mid() { OFS_TUNE="3_1:256:1:1,4:192:1:2,deconv4:128:1:2" timeout 300 python benchmarks/layer_ab.py old_tilings_but_per_phase_deconv2 >> $O 2>> gpurun_out/ab_tilings2.err; }
old; new; mid; old; new; mid
python - <<'PY'
import json
for l in open("gpurun_out/ab_tilings2.jsonl"):
    d = json.loads(l); u = d["us"]
    print(f'{d["tag"]:36s} {d["pairs_s_1"]} {d["pairs_s_2"]} dense {d["dense_ms"]} c1 {u["1"]} 3_1 {u["3_1"]} 4 {u["4"]} d4 {u["deconv4"]} d2 {u["deconv2"]}')
PY
tail -2 gpurun_out/ab_tilings2.err
