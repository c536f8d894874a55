#!/bin/bash
# LLM endpoint bring-up: build, parity tests, small timing run
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_llm.py -x -q -m gpu 2>&1 | tail -40 | tee gpurun_out/pytest_llm.log
echo "pytest rc=${PIPESTATUS[0]}"
if [ -f scripts/llm_bench.py ]; then timeout 900 python scripts/llm_bench.py ${LLM_BENCH_ARGS:-} 2>&1 | tail -30 | tee gpurun_out/llm_bench.log; fi
