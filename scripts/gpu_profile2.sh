#!/bin/bash
# Round-end style pass: full GPU test suite, default bench (both arms), ncu launch list of the bench command, ncu --set
# full captures of the dominant kernels (summaries are copied into profiles/ by scripts/summarise_profiles.py).
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== bench reference arm"; /usr/bin/time -f "wall %es" timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cut -c1-600 gpurun_out/bench_ref.json; tail -1 gpurun_out/bench_ref.err
echo "== bench (default flags)"; /usr/bin/time -f "wall %es" timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('metric','value','unit','n_gpus','steps','warmup','ms_per_step','dtype','gpu_launches','clocks')})
print('e2e', d['e2e']); print('roofline', d['roofline']); print('cpu_baseline', d['cpu_baseline'])
for k,v in d['workloads'].items():
    if v: print(k, {kk: (round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','ms_per_step','error','parity_rel_err_vs_torch_cpu_fp32')}, 'e2e', v.get('e2e',{}).get('value'), (v.get('roofline') or v.get('roofline_prefill') or {}).get('frac'))
PY
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-plugin --no-llama --cpu-seconds 0.2 > gpurun_out/bench_under_ncu.log 2>&1; echo "rc=$?"
for k in forest_staged gemm_tn_pair gemm_tn_persistent attention_varlen layernorm_kernel nchw_to_s2d maxpool3x3s2; do
  echo "== ncu full $k"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 2 -f -o gpurun_out/ncu_$k python bench.py --steps 6 --warmup 3 --no-plugin --no-llama --cpu-seconds 0.1 > gpurun_out/ncu_$k.log 2>&1; echo "rc=$?"
done
ls -la gpurun_out/*.ncu-rep
