#!/bin/bash
# residual-slab L2 prefetch in the GEMM epilogue: A/B on the memory-bound residual GEMMs, op tests, BERT+ResNet bench
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
echo "== residual gemm bench (prefetch ON)"; timeout 300 python scripts/residual_gemm_bench.py 2>&1 | tail -8 | tee gpurun_out/residual_gemm_on.txt
echo "== residual gemm bench (prefetch OFF)"; B2S_RES_PREFETCH=0 timeout 300 python scripts/residual_gemm_bench.py 2>&1 | tail -8 | tee gpurun_out/residual_gemm_off.txt
echo "== ops/bert/resnet tests"; timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_bert.py tests/test_gpu_resnet.py -x -q -m gpu 2>&1 | tail -4
for pf in 1 0; do
echo "== bench (bert + resnet) prefetch=$pf"; B2S_RES_PREFETCH=$pf timeout 900 python bench.py --no-llama --no-plugin --cpu-seconds 0.3 > gpurun_out/bench_br_pf$pf.json 2> gpurun_out/bench_br_pf$pf.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_br_pf$pf.json').read().strip().splitlines()[-1])
for k,v in d['workloads'].items():
    if v: print(k, {kk: (round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','ms_per_step','error')}, v.get('e2e',{}).get('value'), v.get('roofline',{}).get('frac'))
PY
done
