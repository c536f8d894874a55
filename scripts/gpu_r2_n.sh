#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "attention" > gpurun_out/r2n_pytest_attn.log 2>&1; echo "attn pytest rc=$?" >> gpurun_out/r2n_pytest_attn.log
tail -5 gpurun_out/r2n_pytest_attn.log | cut -c1-300
{ echo "=== tcgen05 form"; timeout 300 python scripts/attention_bench.py; echo "=== mma.sync form (B2S_ATTN_TC=0)"; B2S_ATTN_TC=0 timeout 300 python scripts/attention_bench.py; } > gpurun_out/r2n_attention_bench.txt 2>&1
cat gpurun_out/r2n_attention_bench.txt
timeout 300 python scripts/attention_timing.py > gpurun_out/r2n_attn_timing.txt 2>&1; grep -A6 "all256" gpurun_out/r2n_attn_timing.txt | cut -c1-330
timeout 600 python -m pytest tests/test_gpu_bert.py -x -q -m gpu > gpurun_out/r2n_pytest_bert.log 2>&1; echo "bert pytest rc=$?"; tail -3 gpurun_out/r2n_pytest_bert.log
