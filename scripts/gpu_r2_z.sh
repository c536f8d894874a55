#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 --no-plugin --no-ref-path --no-bert --no-resnet --no-llama --cpu-seconds 2 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2z_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z_bench.json').read().strip().splitlines()[-1])
print(d['value'], json.dumps(d['e2e']))
PY
