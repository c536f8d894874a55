#!/bin/bash
# cta_group::2 GEMM (B2S_GEMM_2SM=1): op tests under a watchdog, then throughput against the pair kernel
set -u
mkdir -p gpurun_out
echo "== gemm op tests (2sm)"; B2S_GEMM_2SM=1 timeout 120 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "gemm" --timeout 60 2>&1 | tail -8
echo "== gemm bench (2sm)"; B2S_GEMM_2SM=1 timeout 120 python scripts/gemm_bench.py 2>&1 | tail -9 | cut -c1-175 | tee gpurun_out/gemm_bench_2sm.txt
echo "== gemm bench (pair)"; timeout 120 python scripts/gemm_bench.py 2>&1 | tail -9 | cut -c1-175 | tee gpurun_out/gemm_bench_pair.txt
