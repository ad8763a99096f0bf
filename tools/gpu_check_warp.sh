#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_samplers.py tests/test_gpu_clip.py tests/test_gpu_modes.py -m gpu -q -x --timeout 600 --tb=short 2>&1 | tail -6 > gpurun_out/t_warp.log
cat gpurun_out/t_warp.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_warp.log 2> gpurun_out/bench_warp.err; tail -2 gpurun_out/bench_warp.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_warp.log").read().strip().splitlines()[-1])
print("value", round(d["value"]), "one-at-a-time", round(d["value_one_step_at_a_time"]), "e2e", round(d["e2e"]["value"]), "clip", round(d["e2e_clip_driver"]["value"]))
print("roofline", d["roofline"]["frac"], d["roofline"]["ms_per_step_in_kernel"], "warp", d["roofline_warp"]["frac"], d["roofline_warp"]["ms_per_launch"])
PY
timeout 600 python benchmarks/sampler_sweep.py --quick 2>&1 | grep -v "staged\|direct" | tail -40 > gpurun_out/sweep_quick.txt; cat gpurun_out/sweep_quick.txt
