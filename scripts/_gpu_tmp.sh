#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 2400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r2H_bench_n4.json 2> gpurun_out/r2H_bench_n4.err
echo "bench n4 rc=$?"; tail -5 gpurun_out/r2H_bench_n4.err | cut -c1-300
