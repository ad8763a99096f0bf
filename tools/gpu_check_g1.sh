#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q -x --timeout 600 --tb=short 2>&1 | tail -15 > gpurun_out/t_net.log
tail -8 gpurun_out/t_net.log
timeout 900 python -m pytest tests/test_gpu_samplers.py tests/test_gpu_clip.py -m gpu -q --timeout 600 --tb=short 2>&1 | tail -30 > gpurun_out/t_rest.log
tail -25 gpurun_out/t_rest.log
O=gpurun_out/probe_g1.log
: > $O
run() { timeout 300 python benchmarks/conv_bench.py --batch 8 --iters 20 "$@" >> $O 2>&1; }
run --layers 5,5_1 --variants "256:6:1"
run --layers 6,6_1 --variants "256:8:1"
OFS_FUSED_REDUCE=0 run --layers 5,5_1 --variants "256:6:1"
OFS_FUSED_REDUCE=0 run --layers 6,6_1 --variants "256:8:1"
cat $O | cut -c1-100
for fr in 1 0; do
OFS_FUSED_REDUCE=$fr timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustained-seconds 0 > gpurun_out/bench_g1_$fr.log 2>&1
python - $fr <<'PY'
import json, sys
fr = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_g1_{fr}.log").read().strip().splitlines()[-1])
    print(f"fused_reduce={fr}: value {d['value']:.0f} one-at-a-time {d['value_one_step_at_a_time']:.0f} gemm-set ms {d['roofline']['ms_per_step_in_kernel']:.4f} ach {d['roofline']['achieved']:.1f} frac {d['roofline']['frac']:.3f} launches/step {d['launches_per_step']}")
    print("  " + "  ".join("%s=%.1f" % (b["kernel"].replace("gemm:", ""), b["ms"] * 1e3) for b in d["breakdown"]))
except Exception as e:
    print("failed", e); print(open(f"gpurun_out/bench_g1_{fr}.log").read()[-2500:])
PY
done
