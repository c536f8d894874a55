#!/bin/bash
# Tests + bench + ncu launch list + ncu full captures of the dominant kernels (copied into profiles/ afterwards).
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 --timeout 180 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== bench (memcpy input)"; B2S_ZEROCOPY_IN=0 timeout 600 python bench.py --no-plugin --no-bert --cpu-seconds 1 > gpurun_out/bench_nozcin.json 2> gpurun_out/bench_nozcin.err; echo "rc=$?"; cat gpurun_out/bench_nozcin.json
echo "== bench reference arm"; timeout 300 python bench.py --impl reference --steps 2000 --warmup 10 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cat gpurun_out/bench_ref.json
echo "== gemm bench"; timeout 300 python scripts/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; echo "rc=$?"; cat gpurun_out/gemm_bench.log
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-plugin --cpu-seconds 0.2 > gpurun_out/bench_under_ncu.log 2>&1; echo "rc=$?"
for k in forest_staged gemm_tn_persistent attention_varlen layernorm_kernel embed_layernorm; do
  echo "== ncu full $k"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 2 -f -o gpurun_out/ncu_$k python bench.py --steps 10 --warmup 3 --no-plugin --cpu-seconds 0.2 > gpurun_out/ncu_$k.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_$k.log
done
ls -la gpurun_out | head -40
