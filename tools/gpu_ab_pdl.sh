#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/ab_pdl.jsonl
for m in 0 3 4 0 3 4; do OFS_PDL=$m timeout 300 python benchmarks/layer_ab.py pdl$m >> gpurun_out/ab_pdl.jsonl 2> gpurun_out/ab_pdl.err; done
cut -c1-110 gpurun_out/ab_pdl.jsonl; tail -3 gpurun_out/ab_pdl.err
