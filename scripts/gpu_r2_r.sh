#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "attention" > gpurun_out/r2r_pytest_attn.log 2>&1; echo "attn pytest rc=$?" >> gpurun_out/r2r_pytest_attn.log
tail -5 gpurun_out/r2r_pytest_attn.log | cut -c1-300
{ echo "=== tcgen05 form, every other exp2 on the FMA pipe"; timeout 300 python scripts/attention_bench.py; echo "=== tcgen05 form, all exp2 on MUFU (B2S_ATTN_POLY=0)"; B2S_ATTN_POLY=0 timeout 300 python scripts/attention_bench.py; } > gpurun_out/r2r_attention_bench.txt 2>&1
cat gpurun_out/r2r_attention_bench.txt
timeout 600 python -m pytest tests/test_gpu_bert.py -x -q -m gpu > gpurun_out/r2r_pytest_bert.log 2>&1; echo "bert pytest rc=$?"; tail -3 gpurun_out/r2r_pytest_bert.log
