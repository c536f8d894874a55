#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python bench.py --steps 20 --warmup 3 > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2t_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2t_bench_ref.json 2> gpurun_out/r2t_bench_ref.err; echo "ref rc=$?"
# launch list of the bench (shares), forest traffic capture
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2t_launches.csv python bench.py --steps 3 --warmup 3 --no-plugin --no-ref-path --no-llama --cpu-seconds 1 > gpurun_out/r2t_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:forest_wide -s 5 -c 1 -o gpurun_out/r2_forest_wide python bench.py --steps 3 --warmup 3 --no-plugin --no-ref-path --no-llama --no-bert --no-resnet --cpu-seconds 1 > gpurun_out/r2t_ncu_forest.log 2>&1; echo "ncu forest rc=$?"
