#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/forest_timing.py > gpurun_out/forest_timing.log 2>&1; echo "forest_timing rc=$?"; cat gpurun_out/forest_timing.log
timeout 300 python scripts/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; echo "gemm_bench rc=$?"; cat gpurun_out/gemm_bench.log
