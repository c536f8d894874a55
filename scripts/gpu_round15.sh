#!/bin/bash
# split-K tail of the CTA-pair GEMM: tests under a watchdog, A/B numbers, BERT + ResNet + LLM bench
set -u
mkdir -p gpurun_out
echo "== ops tests"; timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu 2>&1 | tail -12
echo "== bert/resnet/llm tests"; timeout 1200 python -m pytest tests/test_gpu_bert.py tests/test_gpu_resnet.py tests/test_gpu_llm.py -x -q -m gpu 2>&1 | tail -6
echo "== residual gemm bench (split-K ON)"; timeout 300 python scripts/residual_gemm_bench.py 2>&1 | tail -3 | tee gpurun_out/residual_gemm_splitk_on.txt
echo "== residual gemm bench (split-K OFF)"; B2S_SPLIT_K=0 timeout 300 python scripts/residual_gemm_bench.py 2>&1 | tail -3 | tee gpurun_out/residual_gemm_splitk_off.txt
echo "== gemm bench (ON)"; timeout 300 python scripts/gemm_bench.py 2>&1 | tail -9 | cut -c1-200 | tee gpurun_out/gemm_bench_splitk_on.txt
for pf in 1 0; do
echo "== bench (bert + resnet) split_k=$pf"; B2S_SPLIT_K=$pf timeout 900 python bench.py --no-llama --no-plugin --cpu-seconds 0.3 > gpurun_out/bench_br_splitk$pf.json 2> gpurun_out/bench_br_splitk$pf.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_br_splitk$pf.json').read().strip().splitlines()[-1])
for k,v in d['workloads'].items():
    if v: print(k, {kk: (round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','ms_per_step','error','parity_rel_err_vs_torch_cpu_fp32')}, v.get('e2e',{}).get('value'), v.get('roofline',{}).get('frac'))
PY
done
