#!/bin/bash
# full GPU suite + bench + forest ncu traffic
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2s_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_pytest_gpu.log
tail -8 gpurun_out/r2s_pytest_gpu.log | cut -c1-300
{ echo "=== tcgen05 form"; timeout 300 python scripts/attention_bench.py; } > gpurun_out/r2s_attention_bench.txt 2>&1; cat gpurun_out/r2s_attention_bench.txt
timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r2s_bench.err
