#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/clip1_launches.csv python benchmarks/clip_one.py 1 4 > gpurun_out/clip1_ncu.log 2>&1
tail -2 gpurun_out/clip1_ncu.log
