#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_llm.py -x -q -m gpu 2>&1 | tail -4
echo "== skinny 1 CTA/SM"; B2S_SKINNY_CTAS=1 timeout 600 python scripts/llm_bench.py --waves 2 2>&1 | tail -1 | cut -c1-330
echo "== skinny 2 CTA/SM"; timeout 600 python scripts/llm_bench.py --waves 2 2>&1 | tail -1 | cut -c1-330
echo "== bench.py"
timeout 1200 python bench.py --no-bert --no-resnet --no-plugin --cpu-seconds 1 > gpurun_out/bench_llm.json 2> gpurun_out/bench_llm.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_llm.json').read().strip().splitlines()[-1])
print(json.dumps(d['workloads']['llama3_8b'], indent=1)[:3000])
PY
tail -3 gpurun_out/bench_llm.err
