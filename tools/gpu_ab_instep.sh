#!/bin/bash
# in-step A/B of single tiling switches (one step at a time through layer_ab repeats to +-0.1 % on a box)
mkdir -p gpurun_out
O=gpurun_out/ab_instep.jsonl
: > $O
run() { OFS_TUNE="$2" timeout 300 python benchmarks/layer_ab.py "$1" >> $O 2>> gpurun_out/ab_instep.err; }
for rep in 1 2; do
  run base ""
  run deconv3_1cta "deconv3:128:1:1"
  run conv3_1cta "3:256:1:1"
  run conv5_ks4 "5:256:4:1,5_1:256:4:1"
  run conv6_256x8 "6:256:8:1,6_1:256:8:1"
  run deconv5_pairs "deconv5:64:1:2"
  run conv4_1_pairs "4_1:192:1:2"
  run deconv2_pairs "deconv2:64:1:2"
done
python - <<'PY'
import json, collections
acc = collections.OrderedDict()
for l in open("gpurun_out/ab_instep.jsonl"):
    d = json.loads(l)
    acc.setdefault(d["tag"], []).append((d["pairs_s_1"], d["pairs_s_2"]))
for k, v in acc.items():
    print(f"{k:16s}", [x[0] for x in v], [x[1] for x in v])
PY
tail -2 gpurun_out/ab_instep.err
