#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu (parity + engine)"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_engine.py -m gpu -q --maxfail=30 --timeout 180 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu.log
for piece in 4096 8192 32768; do
  echo "== forest timing piece=$piece"; B2S_FOREST_BULK_PIECE=$piece timeout 300 python scripts/forest_timing.py > gpurun_out/forest_timing_$piece.log 2>&1; echo "rc=$?"; sed -n 19,27p gpurun_out/forest_timing_$piece.log
done
echo "== pytest gpu ops (DL kernels)"; timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_bert.py -m gpu -q --maxfail=30 --timeout 180 -p no:cacheprovider > gpurun_out/pytest_ops.log 2>&1; echo "pytest ops rc=$?"; tail -40 gpurun_out/pytest_ops.log
echo "== gemm bench"; timeout 300 python scripts/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; echo "rc=$?"; cat gpurun_out/gemm_bench.log
