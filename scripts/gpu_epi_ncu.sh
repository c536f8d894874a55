#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tn -s 4 -c 2 -f -o gpurun_out/epi_probe python scripts/epi_probe.py > gpurun_out/epi_probe.log 2>&1; echo "rc=$?"
ls -la gpurun_out/epi_probe.ncu-rep
