#!/bin/bash
# 2-GPU check of the bench harness exactly as the driver launches it (torchrun, NCCL).
set -u
mkdir -p gpurun_out
nvidia-smi -L
python -c "import __graft_entry__ as g; g.build()" 
echo "== reference arm x2"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus 2 --steps 200 --warmup 5 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "rc=$?"; cat gpurun_out/bench_ref_n2.json; tail -3 gpurun_out/bench_ref_n2.err
echo "== b200 arm x2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 10 --no-bert --no-resnet --cpu-seconds 1 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"; cat gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
echo "== b200 arm x1 (same flags)"; timeout 600 python bench.py --gpus 1 --steps 200 --warmup 10 --no-bert --no-resnet --cpu-seconds 1 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"; cat gpurun_out/bench_n1.json
echo "== in-process replica router over 2 GPUs"; timeout 300 python - <<'PY'
import asyncio, json, numpy as np, sys
sys.path.insert(0, ".")
from clearml_serving_b200 import BasePreprocessRequest, ModelEndpoint, formats
from oracle import oracle as orc
forest = orc.synth_xgb_forest(200, 6, 32, seed=0)
pm = formats.pack_forest(forest, "xgb", base=0.5)
ep = ModelEndpoint(engine_type="b200", serving_url="rr", auxiliary_cfg={"max_batch_size": 64, "b200.devices": "0,1"})
cls = BasePreprocessRequest.get_engine_cls("b200")
eng = cls.__new__(cls); BasePreprocessRequest.__init__(eng, model_endpoint=ep, task=None); eng._model = pm; eng._b200_setup()
X = np.random.default_rng(0).standard_normal((400, 32)).astype(np.float32)
async def main():
    return await asyncio.gather(*[eng.process(X[i:i+1], {}, None) for i in range(400)])
outs = np.concatenate(asyncio.run(main()))
assert np.array_equal(outs, orc.forest_predict_xgb(forest, X, 0.5)), "router results differ"
print("router ok:", json.dumps(eng.engine_stats()))
eng.unload()
PY
echo "router rc=$?"
