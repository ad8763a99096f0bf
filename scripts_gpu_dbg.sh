#!/bin/bash
for tune in "3_1:256:1:1:16" "3_1:256:1:1:23" "3_1:256:1:1:31" "2:128:1:1:16" "2:128:1:1:31" "1:64:1:1:16"; do
  echo "=== OFS_TUNE=$tune"
  OFS_TUNE="$tune" timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu-baseline 2>&1 | grep "conv dbg" | sort | uniq -c | sort -rn | head -8
done
