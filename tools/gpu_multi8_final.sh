#!/bin/bash
# 8-GPU box, final build of the round: the NCCL gather test, BASELINE configs[3] at N = 8 and 1, the bench line at N = 8
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --timeout 250 --tb=short 2>&1 | tail -3
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29508 \
  benchmarks/config4_clips.py --json gpurun_out/config4_n8.json > gpurun_out/config4_n8.log 2>&1
tail -1 gpurun_out/config4_n8.log | cut -c1-900
timeout 400 python benchmarks/config4_clips.py --json gpurun_out/config4_n1.json > gpurun_out/config4_n1.log 2>&1
tail -1 gpurun_out/config4_n1.log | cut -c1-600
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --gpus 8 --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n8.log 2> gpurun_out/bench_n8.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n8.log").read().strip().splitlines()[-1])
print("N=8 value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "clip", round(d["e2e_clip_driver"]["value"]), "roofline", round(d["roofline"]["frac"], 4), d["clocks"])
PY
