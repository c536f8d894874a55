#!/bin/bash
# One GPU-box visit: smoke, GPU parity tests, bench (both arms), ncu launch list + full capture.
# Usage (from the build container):  gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh'
set -u
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q --maxfail=30 --timeout 180 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== bench (no zero-copy out)"; B2S_ZEROCOPY_OUT=0 timeout 300 python bench.py --no-plugin --cpu-seconds 0.5 > gpurun_out/bench_nozc.json 2> gpurun_out/bench_nozc.err; echo "rc=$?"; cat gpurun_out/bench_nozc.json
echo "== bench reference arm"; timeout 300 python bench.py --impl reference --steps 2000 --warmup 10 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cat gpurun_out/bench_ref.json
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-plugin --cpu-seconds 0.2 > gpurun_out/bench_under_ncu.log 2>&1; echo "rc=$?"
echo "== ncu full"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:forest_pairs -s 5 -c 3 -f -o gpurun_out/forest_pairs python bench.py --steps 20 --warmup 3 --no-plugin --cpu-seconds 0.2 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
