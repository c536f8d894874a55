#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
for ahead in 0 4 8 16 32; do
  echo "== l2_ahead $ahead (old attention)"
  B2S_LLM_ATTN_STREAM=0 B2S_SKINNY_L2_AHEAD=$ahead timeout 600 python scripts/llm_bench.py --waves 3 2>/dev/null | cut -c1-330
done
timeout 300 python -m pytest tests/test_gpu_llm.py -x -q -k "skinny or golden" 2>&1 | tail -3
