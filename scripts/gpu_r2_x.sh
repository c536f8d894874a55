#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_llm.py -x -q > gpurun_out/r2x_pytest_llm.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2x_pytest_llm.log
B2S_LLM_ATTN_TIMING=200 timeout 600 python scripts/llm_bench.py --waves 3 --timing --trace gpurun_out/r2x_llm_trace.txt > gpurun_out/r2x_llm_bench.json 2> gpurun_out/r2x_llm_bench.err
echo "bench rc=$?"
grep -v Warning gpurun_out/r2x_llm_bench.err | grep -A14 "stream form"
head -14 gpurun_out/r2x_llm_trace.txt
cat gpurun_out/r2x_llm_bench.json
