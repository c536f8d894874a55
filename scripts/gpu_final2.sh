#!/bin/bash
# Round-end verification of the committed state: full GPU suite, both bench arms (default flags), launch list of the bench command
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== bench reference arm"; SECONDS=0; timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$? wall ${SECONDS}s"
echo "== bench (default flags)"; SECONDS=0; timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$? wall ${SECONDS}s"; tail -2 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('metric','value','unit','n_gpus','steps','warmup','ms_per_step','dtype','gpu_launches','clocks')})
print('e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
for k,v in d['workloads'].items():
    if v: print(k, {kk: (round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','ms_per_step','error','parity_rel_err_vs_torch_cpu_fp32')}, 'e2e', v.get('e2e',{}).get('value'), {kk: round(vv.get('frac', 0), 3) for kk, vv in v.items() if kk.startswith('roofline') and isinstance(vv, dict)})
PY
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 10 --warmup 3 --no-plugin --no-llama --cpu-seconds 0.1 > gpurun_out/bench_under_ncu.log 2>&1; echo "rc=$?"
