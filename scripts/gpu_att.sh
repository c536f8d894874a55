#!/bin/bash
set -u
echo "== attention tests"; timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_bert.py -x -q -m gpu --timeout 120 -k "attention or bert" 2>&1 | tail -3
echo "== attention bench"; timeout 300 python scripts/attention_bench.py 2>&1 | tail -4
echo "== bench (bert)"; timeout 600 python bench.py --no-llama --no-plugin --no-resnet --cpu-seconds 0.2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for k,v in d['workloads'].items():
    if v: print(' ', k, round(v['value']), round(v['ms_per_step'],3), 'ms', v.get('parity_rel_err_vs_torch_cpu_fp32'))"
