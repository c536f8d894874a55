#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -3 gpurun_out/r2f_pytest.log
{
echo "=== plain loads (default)"; timeout 300 python scripts/forest_wide_timing.py 2>&1 | head -8
echo "=== bulk 8192"; B2S_FOREST_BULK_PIECE=8192 timeout 300 python scripts/forest_wide_timing.py 2>&1 | head -4
echo "=== plain loads, 512 threads"; B2S_FOREST_WIDE_THREADS=512 timeout 300 python scripts/forest_wide_timing.py 2>&1 | head -4
} > gpurun_out/r2f_wide.txt
grep -v "^   cta" gpurun_out/r2f_wide.txt
