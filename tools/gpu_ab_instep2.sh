#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/ab_instep2.jsonl
: > $O
run() { OFS_TUNE="$2" timeout 300 python benchmarks/layer_ab.py "$1" >> $O 2>> gpurun_out/ab_instep2.err; }
for rep in 1 2 3 4; do
  run base41p "4_1:192:1:2"
  run ks4 "4_1:192:1:2,5:256:4:1,5_1:256:4:1"
  run d2pairs "4_1:192:1:2,deconv2:64:1:2"
  run both "4_1:192:1:2,5:256:4:1,5_1:256:4:1,deconv2:64:1:2"
done
python - <<'PY'
import json, collections
acc = collections.OrderedDict()
for l in open("gpurun_out/ab_instep2.jsonl"):
    d = json.loads(l)
    acc.setdefault(d["tag"], []).append((d["pairs_s_1"], d["pairs_s_2"]))
for k, v in acc.items():
    print(f"{k:10s}", [x[0] for x in v], [x[1] for x in v], "mean2", round(sum(x[1] for x in v) / len(v)))
PY
tail -2 gpurun_out/ab_instep2.err
