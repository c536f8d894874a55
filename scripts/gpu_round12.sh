#!/bin/bash
# TMA-store epilogue: tests (ops / bert / resnet / llm), A/B on the memory-bound GEMMs and the BERT shapes, bench
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
echo "== ops tests"; timeout 1200 python -m pytest tests/test_gpu_ops.py -x -q -m gpu 2>&1 | tail -12
echo "== bert/resnet/llm tests"; timeout 1200 python -m pytest tests/test_gpu_bert.py tests/test_gpu_resnet.py tests/test_gpu_llm.py -x -q -m gpu 2>&1 | tail -6
echo "== residual gemm bench (tma store ON)"; timeout 300 python scripts/residual_gemm_bench.py 2>&1 | tail -8 | tee gpurun_out/residual_gemm_tma_on.txt
echo "== residual gemm bench (tma store OFF)"; B2S_TMA_STORE=0 timeout 300 python scripts/residual_gemm_bench.py 2>&1 | tail -8 | tee gpurun_out/residual_gemm_tma_off.txt
echo "== gemm bench (ON)"; timeout 300 python scripts/gemm_bench.py 2>&1 | tail -9 | cut -c1-260 | tee gpurun_out/gemm_bench_tma_on.txt
for pf in 1 0; do
echo "== bench (bert + resnet) tma_store=$pf"; B2S_TMA_STORE=$pf timeout 900 python bench.py --no-llama --no-plugin --cpu-seconds 0.3 > gpurun_out/bench_br_tma$pf.json 2> gpurun_out/bench_br_tma$pf.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_br_tma$pf.json').read().strip().splitlines()[-1])
for k,v in d['workloads'].items():
    if v: print(k, {kk: (round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','ms_per_step','error')}, v.get('e2e',{}).get('value'), v.get('roofline',{}).get('frac'))
PY
done
