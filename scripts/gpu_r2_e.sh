#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_engine.py -x -q -m gpu > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -5 gpurun_out/r2e_pytest.log
{
echo "=== default (threads 1024, C 16)"; timeout 300 python scripts/forest_wide_timing.py 2>&1
echo "=== null mode (launch floor)"; B2S_FOREST_WIDE_NULL=1 timeout 300 python scripts/forest_wide_timing.py 2>&1 | grep "^trees"
} > gpurun_out/r2e_wide.txt
grep -v "^   cta" gpurun_out/r2e_wide.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-bert --no-resnet --no-llama > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2e_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2e_bench_ref.json 2> gpurun_out/r2e_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/r2e_bench_ref.json | head -c 1500
