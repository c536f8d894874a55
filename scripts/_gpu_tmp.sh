#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 2400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2P_bench_n2.json 2> gpurun_out/r2P_bench_n2.err
echo "bench n2 rc=$?"; tail -3 gpurun_out/r2P_bench_n2.err | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_llm.py tests/test_gpu_parity.py -x -q 2>&1 | tail -2
