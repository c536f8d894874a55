#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_engine.py -x -q -m gpu > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -15 gpurun_out/r2d_pytest.log
{
echo "=== default (threads 1024, C 16)"; timeout 300 python scripts/forest_wide_timing.py 2>&1
echo "=== null mode (launch floor)"; B2S_FOREST_WIDE_NULL=1 timeout 300 python scripts/forest_wide_timing.py 2>&1 | grep -v "cta"
echo "=== bulk piece 4096"; B2S_FOREST_BULK_PIECE=4096 timeout 300 python scripts/forest_wide_timing.py 2>&1 | head -4
echo "=== bulk piece 32768"; B2S_FOREST_BULK_PIECE=32768 timeout 300 python scripts/forest_wide_timing.py 2>&1 | head -4
} > gpurun_out/r2d_wide.txt
grep -v "^   cta" gpurun_out/r2d_wide.txt
