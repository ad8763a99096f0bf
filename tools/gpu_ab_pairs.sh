#!/bin/bash
# A/B of the round-2 tilings: parity first, then per-layer timings with single switches turned back
mkdir -p gpurun_out
timeout 300 ./build/l2_fill > gpurun_out/l2_fill.txt 2>&1; cat gpurun_out/l2_fill.txt
timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -q --timeout 300 --tb=short -k "tilings" 2>&1 | tail -15 > gpurun_out/t_tilings.log
tail -8 gpurun_out/t_tilings.log
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q -x --timeout 600 --tb=short -k "not tilings" 2>&1 | tail -15 > gpurun_out/t_net.log
tail -8 gpurun_out/t_net.log
O=gpurun_out/ab_pairs.jsonl
: > $O
run() { timeout 300 python benchmarks/layer_ab.py "$1" >> $O 2>> gpurun_out/ab_pairs.err; }
run new_defaults
OFS_NOSTACK=1 run nostack
OFS_TUNE="deconv2:64:1:1" run stack_1cta
OFS_NOSTACK=1 OFS_TUNE="4:192:1:1,4_1:192:1:1" run nostack_conv4_1cta
OFS_NOSTACK=1 OFS_TUNE="deconv5:128:1:1,deconv4:128:1:1,deconv3:128:1:1,deconv2:64:1:1" run nostack_deconvs_1cta
OFS_NOSTACK=1 OFS_TUNE="deconv5:128:1:1,deconv4:128:1:1" run nostack_deconv54_1cta
OFS_NOSTACK=1 OFS_TUNE="3_1:256:1:2" run nostack_conv3_1_pair
cat $O
tail -5 gpurun_out/ab_pairs.err
