#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_llm.py -x -q > gpurun_out/r2q_pytest_llm.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2q_pytest_llm.log
for cfg in "32 512 128" "4 3000 128" "1 6000 64" "16 1500 64"; do
  set -- $cfg
  for form in 0 1; do
    echo "== batch $1 prompt $2 gen $3 stream_form=$form"
    B2S_LLM_ATTN_STREAM=$form timeout 600 python scripts/llm_bench.py --batch $1 --prompt $2 --gen $3 --waves 3 2>/dev/null | cut -c1-260
  done
done
