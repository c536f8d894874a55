#!/bin/bash
# round 2, call A: forest / engine parity after the wide kernel + C collate, phase timing of the wide kernel
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_engine.py -x -q -m gpu > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
python scripts/forest_wide_timing.py > gpurun_out/r2a_wide16.txt 2>&1
B2S_FOREST_WIDE_C=8 python scripts/forest_wide_timing.py > gpurun_out/r2a_wide8.txt 2>&1
cat gpurun_out/r2a_wide16.txt gpurun_out/r2a_wide8.txt
python bench.py --steps 200 --warmup 20 --no-bert --no-resnet --no-llama > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 1500 gpurun_out/r2a_bench.json
