#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python scripts/llm_bench.py --waves 2 --trace gpurun_out/r2v_llm_trace.txt > gpurun_out/r2v_llm_bench.json 2> gpurun_out/r2v_llm_bench.err
echo "trace rc=$?"
tail -45 gpurun_out/r2v_llm_bench.err
cat gpurun_out/r2v_llm_bench.json
