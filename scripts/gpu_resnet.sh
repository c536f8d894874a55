#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "== pytest resnet"; timeout 900 python -m pytest tests/test_gpu_resnet.py tests/test_gpu_ops.py -m gpu -q --maxfail=30 --timeout 300 -p no:cacheprovider > gpurun_out/pytest_resnet.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_resnet.log
echo "== bench (resnet+bert)"; timeout 900 python -X faulthandler bench.py --no-plugin --cpu-seconds 4 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(json.dumps(d['workloads'], indent=0))"; tail -20 gpurun_out/bench.err
