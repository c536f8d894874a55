#!/bin/bash
# 2-GPU box: tensor-parallel pair parity (vs single GPU) and Llama-3-8B TP=2 wave timing
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
nvidia-smi topo -m 2>&1 | head -8
timeout 600 python -m pytest tests/test_gpu_llm.py -x -q -m gpu -k tensor_parallel 2>&1 | tail -15 | tee gpurun_out/pytest_llm_tp2.log
echo "== TP2 bench"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 scripts/llm_bench.py --waves 3 2>&1 | grep -v "^W\|OMP_NUM" | tail -5 | tee gpurun_out/llm_bench_tp2.log
echo "== TP1 bench (same box)"
timeout 600 python scripts/llm_bench.py --waves 3 2>&1 | tail -2 | tee gpurun_out/llm_bench_tp1.log
