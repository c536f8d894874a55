#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "== pytest ops+bert+resnet"; timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_bert.py tests/test_gpu_resnet.py -m gpu -q --maxfail=30 --timeout 300 -p no:cacheprovider > gpurun_out/pytest_ops.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_ops.log
echo "== gemm bench (pair)"; timeout 300 python scripts/gemm_bench.py > gpurun_out/gemm_bench_pair.log 2>&1; echo "rc=$?"; cat gpurun_out/gemm_bench_pair.log
echo "== gemm bench (no pair)"; B2S_GEMM_PAIR=0 timeout 300 python scripts/gemm_bench.py > gpurun_out/gemm_bench_nopair.log 2>&1; echo "rc=$?"; cat gpurun_out/gemm_bench_nopair.log
echo "== bench"; timeout 900 python -X faulthandler bench.py --no-plugin --cpu-seconds 2 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); w=d['workloads']; print({k:(v.get('value'),v.get('ms_per_step'),v.get('roofline',{}).get('frac'),v.get('parity_rel_err_vs_torch_cpu_fp32')) for k,v in w.items()})"; tail -20 gpurun_out/bench.err
