#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_llm.py -x -q -m gpu 2>&1 | tail -5
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_llm.csv python scripts/llm_bench.py --layers 4 --waves 1 --gen 5 --no-graph > gpurun_out/llm_under_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, io
from collections import defaultdict
lines=[l for l in open('gpurun_out/launches_llm.csv') if l.startswith('"')]
rows=list(csv.DictReader(io.StringIO("".join(lines))))
agg=defaultdict(list)
for r in rows:
    try: agg[(r["Kernel Name"].split("(")[0][:48], r["Grid Size"], r["Block Size"])].append(float(r["Metric Value"].replace(",","")))
    except Exception: pass
tot=sum(sum(v) for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1]))[:40]:
    print("%-50s grid=%-14s blk=%-12s n=%4d avg=%9.0f ns share=%5.1f%%"%(k[0],k[1],k[2],len(v),sum(v)/len(v),100*sum(v)/tot))
PY
