#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
for sk in 0 1 2 4 6 7 8 16 32 64 128 120 127 255; do
  echo -n "skip=$sk "; B2S_LLM_SKIP=$sk timeout 300 python scripts/llm_bench.py --waves 2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['decode_step_ms'], d['prefill_ms'])"
done 2>&1 | tee gpurun_out/llm_skip.log
