#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q -x --timeout 600 --tb=short 2>&1 | tail -15 > gpurun_out/t_net.log
tail -8 gpurun_out/t_net.log
O=gpurun_out/probe_c1.log
: > $O
run() { timeout 300 python benchmarks/conv_bench.py --batch 8 --iters 20 "$@" >> $O 2>&1; }
run --layers 1 --variants "64:1:4"
run --layers 2 --variants "128:1:4"
run --layers 3 --variants "256:1:2,256:1:1"
run --layers 3_1 --variants "256:1:1"
run --layers 4,4_1 --variants "192:1:1,256:1:2"
run --layers 5,5_1 --variants "256:6:1"
run --layers 6,6_1 --variants "256:8:1"
run --layers deconv5 --variants "64:1:1,64:1:8,128:1:1"
run --layers deconv4,deconv3 --variants "128:1:1,128:1:8"
run --layers deconv2 --variants "64:1:1,64:1:8"
run --layers predict2 --variants "32:1:1,32:1:32"
cat $O | cut -c1-100
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c1.log 2>&1
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_c1.log").read().strip().splitlines()[-1])
    print(f"value {d['value']:.0f} one-at-a-time {d['value_one_step_at_a_time']:.0f} gemm-set ms {d['roofline']['ms_per_step_in_kernel']:.4f} ach {d['roofline']['achieved']:.1f}")
    print("  " + "  ".join("%s=%.1f" % (b["kernel"].replace("gemm:", ""), b["ms"] * 1e3) for b in d["breakdown"]))
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_c1.log").read()[-2500:])
PY
