#!/bin/bash
# cta_group::2 GEMM on the workloads: full op / model tests with it on, then BERT + ResNet + Llama prefill with it on and off
set -u
mkdir -p gpurun_out
echo "== tests (2sm on)"; B2S_GEMM_2SM=1 timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_bert.py tests/test_gpu_resnet.py tests/test_gpu_llm.py -x -q -m gpu --timeout 120 2>&1 | tail -4
for v in 1 0; do
echo "== bench (bert + resnet) 2sm=$v"; B2S_GEMM_2SM=$v timeout 600 python bench.py --no-llama --no-plugin --cpu-seconds 0.2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for k,v in d['workloads'].items():
    if v: print(' ', k, round(v['value']), round(v['ms_per_step'],3), 'ms')"
echo "== llm bench 2sm=$v"; B2S_GEMM_2SM=$v timeout 600 python scripts/llm_bench.py --waves 2 2>&1 | tail -1 | cut -c1-330
done
