#!/bin/bash
# implicit-GEMM convolution: op parity, ResNet parity, ResNet-50 bench implicit vs explicit
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
echo "== conv op + resnet tests"; timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_resnet.py -q -m gpu --maxfail=40 -k "conv or resnet" 2>&1 | tail -40 | tee gpurun_out/pytest_conv.log
for impl in 1 0; do
  echo "== resnet bench implicit=$impl"
  B2S_CONV_IMPLICIT=$impl timeout 600 python bench.py --no-bert --no-llama --no-plugin --cpu-seconds 0.3 --steps 100 --warmup 5 > gpurun_out/bench_conv_$impl.json 2> gpurun_out/bench_conv_$impl.err; echo "rc=$?"; tail -2 gpurun_out/bench_conv_$impl.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_conv_$impl.json').read().strip().splitlines()[-1])
r=d['workloads']['resnet50']
print({k:r.get(k) for k in ('value','ms_per_step','parity_rel_err_vs_torch_cpu_fp32','gpu_launches_per_step','error')}, r.get('e2e'), r.get('roofline',{}).get('frac'))
PY
done
