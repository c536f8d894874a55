#!/bin/bash
# ncu --set full captures of the LLM endpoint's kernels (slim command: 2 layers, 1 wave, 3 decode steps, eager launches)
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
CMD="python scripts/llm_bench.py --layers 2 --waves 1 --gen 4 --no-graph"
cap() { # name regex skip count
  echo "== ncu full $1"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o gpurun_out/ncu_$1 $CMD > gpurun_out/ncu_$1.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/ncu_$1.log | cut -c1-200
}
cap skinny_gemm skinny_gemm 3 2        # gate/up and down projection of layer 0, first decode step
cap llm_attn_decode llm_attn_decode 0 1
cap llm_attn_prefill llm_attn_prefill 0 1
cap llm_reduce_rms llm_reduce_rms 0 1   # prefill form (bf16 partial, 16384 tokens)
cap gemm_tn_pair gemm_tn_pair 2 2       # gate/up and down projection GEMMs of the prefill, layer 0
ls -la gpurun_out/*.ncu-rep | tail -8
