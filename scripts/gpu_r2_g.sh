#!/bin/bash
mkdir -p gpurun_out
run() { echo "=== $1"; env $1 timeout 300 python bench.py --steps 20 --warmup 3 --no-bert --no-resnet --no-llama --no-plugin --no-ref-path --cpu-seconds 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f (%.2f us)  warm %.0f (%.2f us)  e2e %.0f (%.2f us, serial %.2f us)' % (d['value'], d['ms_per_step']*1e3, d['value_l2_warm'], d['ms_per_step_l2_warm']*1e3, d['e2e']['value'], d['e2e']['ms_per_step']*1e3, d['e2e']['ms_per_step_serial']*1e3))"; }
{
run "B2S_X=0"
run "B2S_FOREST_WIDE_R=32"
run "B2S_FOREST_WIDE_C=8"
run "B2S_FOREST_WIDE_C=8 B2S_FOREST_WIDE_R=32"
run "B2S_FOREST_WIDE_THREADS=512"
run "B2S_FOREST_WIDE=0"
run "B2S_ZEROCOPY_IN=0"
run "B2S_ZEROCOPY_IN=0 B2S_ZEROCOPY_OUT=0"
} > gpurun_out/r2g_e2e_variants.txt 2>&1
cat gpurun_out/r2g_e2e_variants.txt
