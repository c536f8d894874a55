#!/bin/bash
# A/B of two attention.cu versions on the same box
set -u
for v in new old new old; do
  cp scripts/tmp_ab/attention_$v.cu.txt clearml_serving_b200/csrc/attention.cu
  python -m clearml_serving_b200.build --force > /dev/null 2>&1
  echo "== $v"; timeout 300 python scripts/attention_bench.py 2>&1 | tail -4 | head -2
  timeout 600 python bench.py --no-llama --no-plugin --no-resnet --cpu-seconds 0.1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
v=d['workloads']['bert_base']; print('  bert', round(v['value']), round(v['ms_per_step'],3), 'ms')"
done
