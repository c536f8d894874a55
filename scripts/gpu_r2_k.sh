#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "attention" > gpurun_out/r2l_pytest_attn.log 2>&1; echo "attn pytest rc=$?" >> gpurun_out/r2l_pytest_attn.log
tail -12 gpurun_out/r2l_pytest_attn.log | cut -c1-300
{ echo "=== tcgen05 form"; timeout 300 python scripts/attention_bench.py; echo "=== mma.sync form (B2S_ATTN_TC=0)"; B2S_ATTN_TC=0 timeout 300 python scripts/attention_bench.py; } > gpurun_out/r2l_attention_bench.txt 2>&1
cat gpurun_out/r2l_attention_bench.txt
timeout 1200 python -m pytest tests/test_gpu_bert.py -x -q -m gpu > gpurun_out/r2l_pytest_llm_bert.log 2>&1; echo "llm+bert pytest rc=$?" >> gpurun_out/r2l_pytest_llm_bert.log
tail -12 gpurun_out/r2l_pytest_llm_bert.log | cut -c1-300
