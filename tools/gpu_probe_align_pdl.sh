#!/bin/bash
# (1) 128-byte aligned concat strides vs the current odd strides, layer by layer; (2) PDL modes with the late trigger
mkdir -p gpurun_out
O=gpurun_out/probe_align.log
: > $O
run() { timeout 300 python benchmarks/conv_bench.py --batch 8 --iters 20 "$@" >> $O 2>&1; }
run --layers 2,2a --variants "128:1:4"
run --layers 3,3a --variants "256:1:2,256:1:1"
run --layers 3_1,3_1a --variants "256:1:1"
run --layers 4,4a --variants "192:1:1"
run --layers 4_1,4_1a --variants "192:1:1"
run --layers 5,5a,5_1,5_1a --variants "256:6:1"
run --layers 6,6a --variants "256:8:1"
run --layers deconv5,deconv5a --variants "64:1:1"
run --layers deconv4,deconv4a,deconv3,deconv3a --variants "128:1:1"
run --layers deconv2,deconv2a --variants "64:1:1"
cat $O | cut -c1-100
for m in 0 1 2; do
  OFS_PDL=$m timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_pdl$m.log 2>&1
  python - $m <<'PY'
import json, sys
m = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_pdl{m}.log").read().strip().splitlines()[-1])
    print(f"OFS_PDL={m}: value {d['value']:.0f} one-at-a-time {d['value_one_step_at_a_time']:.0f} gemm-set ms {d['roofline']['ms_per_step_in_kernel']:.4f} ach {d['roofline']['achieved']:.1f}")
except Exception as e:
    print("failed", m, e); print(open(f"gpurun_out/bench_pdl{m}.log").read()[-1500:])
PY
done
