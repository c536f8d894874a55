#!/bin/bash
# residual loads hoisted above the TMEM read: op tests + A/B numbers
set -u
mkdir -p gpurun_out
echo "== ops tests"; timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_resnet.py -x -q -m gpu 2>&1 | tail -4
echo "== residual gemm bench"; timeout 300 python scripts/residual_gemm_bench.py 2>&1 | tail -8 | tee gpurun_out/residual_gemm_hoist.txt
echo "== bench (resnet)"; timeout 900 python bench.py --no-llama --no-plugin --no-bert --cpu-seconds 0.3 > gpurun_out/bench_r.json 2> gpurun_out/bench_r.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r.json').read().strip().splitlines()[-1])
for k,v in d['workloads'].items():
    if v: print(k, {kk: (round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','ms_per_step','error')}, v.get('e2e',{}).get('value'), v.get('roofline',{}).get('frac'))
PY
