#!/bin/bash
# forest staged kernel, row-owner form: parity (bit-exact) + engine tests under a watchdog, phase timing, headline bench
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
echo "== pytest gpu (parity + engine)"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_engine.py -m gpu -x -q --timeout 120 -p no:cacheprovider 2>&1 | tail -6
echo "== forest timing"; timeout 300 python scripts/forest_timing.py > gpurun_out/forest_timing.log 2>&1; echo "rc=$?"; head -16 gpurun_out/forest_timing.log | cut -c1-230
echo "== bench (forest only)"; timeout 600 python bench.py --cpu-seconds 1 --no-plugin --no-bert --no-resnet --no-llama > gpurun_out/bench_forest.json 2> gpurun_out/bench_forest.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_forest.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms_per_step', d['ms_per_step'], 'warm', round(d['value_l2_warm']), 'e2e', round(d['e2e']['value']), 'serial us', d['e2e']['ms_per_step_serial']*1e3, 'roofline', d['roofline']['frac'])
PY
