#!/bin/bash
# pooled collate + pipelined H2D in b2s_infer_batch: engine / parity / resnet tests, ResNet + BERT + forest e2e
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
echo "== engine/parity/resnet tests"; timeout 1200 python -m pytest tests/test_gpu_engine.py tests/test_gpu_parity.py tests/test_gpu_resnet.py -x -q -m gpu 2>&1 | tail -5
echo "== bench"; timeout 1200 python bench.py --no-llama --cpu-seconds 1 > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_e2e.json').read().strip().splitlines()[-1])
print('forest', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e'].get('ms_per_step_serial'), 'plugin', d.get('plugin',{}).get('closed_loop_req_s'), d.get('plugin',{}).get('poisson'))
for k,v in d['workloads'].items():
    if v: print(k, {kk: (round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','ms_per_step','error')}, 'e2e', v.get('e2e',{}).get('value'), v.get('e2e',{}).get('ms_per_step'), v.get('roofline',{}).get('frac'))
PY
for t in 1 4 8 16; do echo "gather threads $t"; B2S_GATHER_THREADS=$t timeout 600 python bench.py --no-llama --no-plugin --no-bert --cpu-seconds 0.2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); v=d['workloads']['resnet50']; print(' resnet e2e', round(v['e2e']['value']), 'img/s', round(v['e2e']['ms_per_step'],2), 'ms/step')"; done
