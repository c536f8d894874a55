#!/bin/bash
# space-to-depth stem (overlapping-stride 5-D TMA view): op tests, ResNet tests, bench, ResNet launch list
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
echo "== stem op tests"; timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "stem" 2>&1 | tail -15
echo "== ops/resnet tests"; timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_resnet.py -x -q -m gpu 2>&1 | tail -6
echo "== bench (resnet)"; timeout 900 python bench.py --no-llama --no-plugin --no-bert --cpu-seconds 0.3 > gpurun_out/bench_r.json 2> gpurun_out/bench_r.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r.json').read().strip().splitlines()[-1])
for k,v in d['workloads'].items():
    if v: print(k, {kk: (round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','ms_per_step','error','parity_rel_err_vs_torch_cpu_fp32')}, v.get('e2e',{}).get('value'), v.get('roofline',{}).get('frac'))
PY
echo "== ncu launch list (resnet only)"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_resnet.csv python bench.py --steps 3 --warmup 3 --no-plugin --no-bert --no-llama --cpu-seconds 0.1 > gpurun_out/bench_under_ncu_resnet.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv, io
lines=[l for l in open('gpurun_out/launches_resnet.csv') if l.startswith('"')]
rows=list(csv.DictReader(io.StringIO("".join(lines))))
idx=[i for i,r in enumerate(rows) if 'nchw_to_' in r["Kernel Name"]]
if idx:
    seg=rows[idx[-1]:idx[-1]+59]
    tot=0.0
    for r in seg:
        t=float(r["Metric Value"].replace(",","")); tot+=t
        print("%-44s grid=%-14s %9.0f ns"%(r["Kernel Name"].split("(")[0][:44], r["Grid Size"], t))
    print("sum of step (cold, serialised):", tot/1e6, "ms")
PY
