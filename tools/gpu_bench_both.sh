#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; tail -3 gpurun_out/bench_default.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; tail -c 1500 gpurun_out/bench_ref.log; tail -3 gpurun_out/bench_ref.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_default.log").read().strip().splitlines()[-1])
print("value", round(d["value"]), "1-at-a-time", round(d["value_one_step_at_a_time"]), "sustained", round(d["sustained"]["value"]), "e2e", round(d["e2e"]["value"]), "clip", round(d["e2e_clip_driver"]["value"]))
r = d["roofline"]; w = d["roofline_warp"]
print("roofline frac", round(r["frac"], 4), r["ms_per_step_in_kernel"], "hot", r["after_sustained_load"], "share", r["share_of_step"])
print("warp frac", round(w["frac"], 4), w["ms_per_launch"], "hot", w["after_sustained_load"])
print("clocks", d["clocks"])
PY
