#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_llm.py -x -q > gpurun_out/r2O_pytest_llm.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2O_pytest_llm.log
B2S_LLM_ATTN_TIMING=200 timeout 600 python scripts/llm_bench.py --waves 3 --timing --trace gpurun_out/r2O_trace.txt 2> gpurun_out/r2O.err | cut -c1-260
grep -A13 "stream form" gpurun_out/r2O.err | head -15
sed -n 3,6p gpurun_out/r2O_trace.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/llm_bench.py --waves 3 2>/dev/null | cut -c1-260
