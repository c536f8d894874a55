#!/bin/bash
# what the driver runs at round end: the reference arm, then the default bench line
set -u
mkdir -p gpurun_out
echo "== bench reference arm"; SECONDS=0; timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$? wall ${SECONDS}s"; cut -c1-700 gpurun_out/bench_ref.json
echo "== bench (default flags)"; SECONDS=0; timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$? wall ${SECONDS}s"; tail -2 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('metric','value','unit','n_gpus','steps','warmup','ms_per_step','dtype','gpu_launches','clocks')})
print('e2e', d['e2e']); print('roofline', d['roofline']); print('cpu_baseline', d['cpu_baseline'])
for k,v in d['workloads'].items():
    if v: print(k, {kk: (round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','ms_per_step','error','parity_rel_err_vs_torch_cpu_fp32')}, 'e2e', v.get('e2e',{}).get('value'), (v.get('roofline') or v.get('roofline_prefill') or {}).get('frac'))
PY
