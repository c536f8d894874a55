#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_llm.py -x -q -m gpu > gpurun_out/r2h_pytest_llm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest_llm.log
tail -25 gpurun_out/r2h_pytest_llm.log
timeout 300 python -c "
import __graft_entry__ as g
g.smoke()
" > gpurun_out/r2h_smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/r2h_smoke.log
