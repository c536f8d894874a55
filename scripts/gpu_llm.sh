#!/bin/bash
# LLM endpoint: build, parity tests, Llama-3-8B wave timing (TP 1), with and without programmatic dependent launch
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_llm.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_llm.log
echo "pytest rc=${PIPESTATUS[0]}"
echo "== PDL on"; timeout 900 python scripts/llm_bench.py --timing ${LLM_BENCH_ARGS:-} 2>&1 | tail -18 | tee gpurun_out/llm_bench.log
echo "== PDL off"; B2S_LLM_PDL=0 B2S_SKINNY_PDL=0 timeout 900 python scripts/llm_bench.py ${LLM_BENCH_ARGS:-} 2>&1 | tail -3 | tee gpurun_out/llm_bench_nopdl.log
