#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu (parity + engine)"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_engine.py -m gpu -q --maxfail=30 --timeout 180 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu.log
echo "== forest timing"; timeout 300 python scripts/forest_timing.py > gpurun_out/forest_timing.log 2>&1; echo "rc=$?"; head -24 gpurun_out/forest_timing.log
echo "== bench (no bert)"; timeout 600 python -X faulthandler bench.py --cpu-seconds 3 --no-plugin --no-bert > gpurun_out/bench_nobert.json 2> gpurun_out/bench_nobert.err; echo "bench rc=$?"; cat gpurun_out/bench_nobert.json; tail -25 gpurun_out/bench_nobert.err
echo "== bench (bert)"; timeout 600 python -X faulthandler bench.py --cpu-seconds 3 --no-plugin > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -40 gpurun_out/bench.err
