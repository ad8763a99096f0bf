#!/bin/bash
# GPU check: tests (separate processes), smoke, bench, optional ncu passes ($1 = "ncu" to profile)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_samplers.py -m gpu -q --timeout 300 --tb=short 2>&1 | tail -25 > gpurun_out/t_samplers.log
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q --timeout 300 --tb=short -k "conv_gemm" 2>&1 | tail -40 > gpurun_out/t_conv.log
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q --timeout 600 --tb=short -s -k "not conv_gemm" 2>&1 | tail -60 > gpurun_out/t_net.log
timeout 900 python -m pytest tests/test_gpu_clip.py -m gpu -q --timeout 300 --tb=short 2>&1 | tail -25 > gpurun_out/t_clip.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --steps 200 --warmup 5 > gpurun_out/bench.log 2>&1
for f in t_samplers.log t_conv.log t_net.log t_clip.log smoke.log; do echo "=== $f"; tail -30 gpurun_out/$f; done
echo "=== bench"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")})
    print("e2e", d["e2e"]["value"], "roofline", d["roofline"]["achieved"], d["roofline"]["frac"], "warp", d["roofline_warp"]["achieved"], d["roofline_warp"]["frac"])
    print("cpu", d["cpu_baseline"])
    for b in d["breakdown"]:
        print("  %-28s %8.4f ms  %s" % (b["kernel"], b["ms"], ("%.0f TF" % b["tflops"]) if b["tflops"] else ""))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench.log").read()[-3000:])
PY
if [ "$1" == "ncu" ]; then
  CMD="env OFS_GRAPH=0 python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
  $CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
  $CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 30 -c 16 -o gpurun_out/prof_conv -f $CMD > gpurun_out/ncu_conv.log 2>&1
  $CMD > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"warp5|pack_act|predict2_gather|pyr_kernel|splitk" -s 12 -c 12 -o gpurun_out/prof_misc -f $CMD > gpurun_out/ncu_misc.log 2>&1
  tail -3 gpurun_out/ncu_launches.log gpurun_out/ncu_conv.log gpurun_out/ncu_misc.log
  ls -la gpurun_out
fi
if [ "$1" == "sweep" ] || [ "$2" == "sweep" ]; then
  timeout 1200 python benchmarks/sampler_sweep.py --quick > gpurun_out/sweep.log 2>&1; tail -60 gpurun_out/sweep.log
fi
