#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_llm.py -x -q > gpurun_out/r2w_pytest_llm.log 2>&1
echo "pytest rc=$?"
tail -15 gpurun_out/r2w_pytest_llm.log
timeout 600 python scripts/llm_bench.py --waves 3 --trace gpurun_out/r2w_llm_trace.txt > gpurun_out/r2w_llm_bench.json 2> gpurun_out/r2w_llm_bench.err
echo "bench rc=$?"
head -16 gpurun_out/r2w_llm_trace.txt
cat gpurun_out/r2w_llm_bench.json
