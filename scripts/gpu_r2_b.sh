#!/bin/bash
mkdir -p gpurun_out
for cfg in "1024 16" "512 16" "256 16" "1024 8" "512 8" "1024 4"; do
  set -- $cfg
  echo "=== threads=$1 C=$2"
  B2S_FOREST_WIDE_THREADS=$1 B2S_FOREST_WIDE_C=$2 python scripts/forest_wide_timing.py 2>&1
done > gpurun_out/r2b_wide_variants.txt
cat gpurun_out/r2b_wide_variants.txt
