#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_engine.py -x -q -m gpu > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -15 gpurun_out/r2c_pytest.log
for cfg in "1024 16" "1024 8"; do
  set -- $cfg
  echo "=== threads=$1 C=$2"
  B2S_FOREST_WIDE_THREADS=$1 B2S_FOREST_WIDE_C=$2 timeout 300 python scripts/forest_wide_timing.py 2>&1
done > gpurun_out/r2c_wide_variants.txt
grep -v "^   cta" gpurun_out/r2c_wide_variants.txt
