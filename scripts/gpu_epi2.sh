#!/bin/bash
set -u
for d in 0 2 6 10 14 18 30; do B2S_EPI_DBG=$d timeout 120 python scripts/epi_probe2.py 2>&1 | grep dbg; done
