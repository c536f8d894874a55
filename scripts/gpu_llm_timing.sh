#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
echo "== PDL on"; timeout 900 python scripts/llm_bench.py --waves 2 --timing 2>&1 | tail -22 | tee gpurun_out/llm_timing.log
echo "== PDL off"; B2S_SKINNY_PDL=0 timeout 900 python scripts/llm_bench.py --waves 2 --timing 2>&1 | tail -22 | tee gpurun_out/llm_timing_nopdl.log
