#!/bin/bash
mkdir -p gpurun_out
timeout 100 python benchmarks/clip_one.py 8 3 > gpurun_out/clip_plain.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"warp5_u8|assemble_x0|resize_u8" -s 4 -c 4 -o gpurun_out/prof_clip -f python benchmarks/clip_one.py 8 3 > gpurun_out/ncu_clip.log 2>&1
tail -2 gpurun_out/ncu_clip.log; ls -la gpurun_out/prof_clip.ncu-rep
