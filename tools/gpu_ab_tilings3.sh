#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err; python - $tag <<'PY'
import json, sys
t = sys.argv[1]
d = json.loads(open(f"gpurun_out/bench_{t}.log").read().strip().splitlines()[-1])
print(f'{t:10s} value {d["value"]:.0f} one-at-a-time {d["value_one_step_at_a_time"]:.0f} sustained {d["sustained"]["value"]:.0f} clip {d["e2e_clip_driver"]["value"]:.0f} roofline {d["roofline"]["frac"]:.4f} ({d["roofline"]["ms_per_step_in_kernel"]:.4f} ms) hot {d["roofline"]["after_sustained_load"]["frac_of_burst_peak"]:.4f}')
PY
}
run old1 OFS_STACK=1 OFS_TUNE="3_1:256:1:1,4:192:1:2,deconv4:128:1:2"
run new1 OFS_X=0
run mid1 OFS_TUNE="3_1:256:1:1,4:192:1:2,deconv4:128:1:2"
run old2 OFS_STACK=1 OFS_TUNE="3_1:256:1:1,4:192:1:2,deconv4:128:1:2"
run new2 OFS_X=0
run mid2 OFS_TUNE="3_1:256:1:1,4:192:1:2,deconv4:128:1:2"
