#!/bin/bash
# attention v2 ((sequence, head) CTAs, K/V resident in smem): tests, microbench, BERT bench
set -u
mkdir -p gpurun_out
echo "== attention tests"; timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_bert.py -x -q -m gpu 2>&1 | tail -4
echo "== attention bench"; timeout 300 python scripts/attention_bench.py 2>&1 | tail -4 | tee gpurun_out/attention_bench.txt
echo "== bench (bert)"; timeout 900 python bench.py --no-llama --no-plugin --no-resnet --cpu-seconds 0.3 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_b.json').read().strip().splitlines()[-1])
for k,v in d['workloads'].items():
    if v: print(k, {kk: (round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','ms_per_step','error')}, v.get('e2e',{}).get('value'), v.get('roofline',{}).get('frac'))
PY
