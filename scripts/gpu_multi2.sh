#!/bin/bash
# 2-GPU box: TP-2 parity test, the new collate test, the default bench under torchrun (what the driver's scaling run does)
set -u
mkdir -p gpurun_out
nvidia-smi -L | head -4
echo "== tests (tp2 + collate + llm streaming)"; timeout 900 python -m pytest tests/test_gpu_llm.py tests/test_gpu_resnet.py -x -q -m gpu 2>&1 | tail -4
echo "== bench N=2 (torchrun)"; SECONDS=0
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$? wall ${SECONDS}s"; tail -3 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_n2.json').read().strip().splitlines() if l.startswith('{')][-1])
print({k: d.get(k) for k in ('metric','value','n_gpus','ms_per_step','scaling','gpu_launches')})
print('e2e', d['e2e']['value']); print('plugin', {k: v for k, v in d.get('plugin', {}).items() if k.startswith('poisson')})
for k,v in d['workloads'].items():
    if v: print(k, {kk: (round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','ms_per_step','error','config')}, 'e2e', v.get('e2e',{}).get('value'))
PY
echo "== reference arm N=2"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 2>/dev/null | cut -c1-300
