#!/bin/bash
# 2 GPUs: TP=2 parity test, bench at N=2 (replicas + tensor-parallel pair), vLLM TP=2 comparator
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_llm.py -x -q -m gpu -k "tensor_parallel" > gpurun_out/r2u_pytest_tp.log 2>&1; echo "tp pytest rc=$?"; tail -3 gpurun_out/r2u_pytest_tp.log | cut -c1-300
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2u_bench_n2.json 2> gpurun_out/r2u_bench_n2.err; echo "bench n2 rc=$?"; tail -c 500 gpurun_out/r2u_bench_n2.err
timeout 1200 python scripts/vllm_compare.py --tp 2 --waves 3 --out gpurun_out/r2_vllm_tp2.json > gpurun_out/r2_vllm_tp2.log 2>&1; echo "vllm tp2 rc=$?"; cat gpurun_out/r2_vllm_tp2.json
