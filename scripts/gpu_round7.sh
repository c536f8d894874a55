#!/bin/bash
# after: grouped GEMM tile walk, SwiGLU epilogue, persistent N=64 GEMM
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
echo "== llm bench"; timeout 600 python scripts/llm_bench.py --waves 3 2>&1 | tail -1 | cut -c1-400
echo "== llm bench slab=1024MB (old order)"; B2S_GEMM_SLAB_MB=4096 timeout 600 python scripts/llm_bench.py --waves 2 2>&1 | tail -1 | cut -c1-250
echo "== gemm bench"; timeout 300 python scripts/gemm_bench.py 2>&1 | tail -9 | cut -c1-260
echo "== bench.py"; timeout 1500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
w=d.pop('workloads'); 
print({k:d[k] for k in ('value','ms_per_step','value_l2_warm','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'])
for k,v in w.items():
    if v is None: print(k, None); continue
    print(k, {kk: (round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','ms_per_step','prefill_ms','decode_step_ms','error','parity_rel_err','roofline')})
PY
