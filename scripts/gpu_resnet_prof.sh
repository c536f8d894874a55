#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
echo "== ncu launch list (resnet only)"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_resnet.csv python bench.py --steps 5 --warmup 3 --no-plugin --no-bert --cpu-seconds 0.2 > gpurun_out/bench_under_ncu_resnet.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv, io
from collections import defaultdict
lines=[l for l in open('gpurun_out/launches_resnet.csv') if l.startswith('"')]
rows=list(csv.DictReader(io.StringIO("".join(lines))))
agg=defaultdict(list)
for r in rows:
    try: agg[(r["Kernel Name"].split("(")[0][:60], r["Grid Size"])].append(float(r["Metric Value"].replace(",","")))
    except Exception: pass
tot=sum(sum(v) for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1]))[:40]:
    print("%-62s grid=%-16s n=%4d avg=%9.0f ns share=%5.1f%%"%(k[0],k[1],len(v),sum(v)/len(v),100*sum(v)/tot))
PY
