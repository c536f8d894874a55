#!/bin/bash
# Quick GPU visit: smoke + GPU tests + short bench (no ncu).  gpurun --timeout 900 -- 'bash scripts/gpu_quick.sh'
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu (parity + engine)"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_engine.py -m gpu -q --maxfail=30 --timeout 180 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py --cpu-seconds 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== bench reference arm"; timeout 300 python bench.py --impl reference --steps 2000 --warmup 10 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cat gpurun_out/bench_ref.json
echo "== pytest gpu ops (DL kernels)"; timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_bert.py -m gpu -q --maxfail=30 --timeout 120 -p no:cacheprovider > gpurun_out/pytest_ops.log 2>&1; echo "pytest ops rc=$?"; tail -60 gpurun_out/pytest_ops.log
