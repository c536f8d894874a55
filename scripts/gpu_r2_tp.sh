#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_llm.py -x -q -k "tensor_parallel or golden or mixed" 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/llm_bench.py --waves 3 --trace gpurun_out/r2tp_trace_rank.txt > gpurun_out/r2tp_bench.json 2> gpurun_out/r2tp_bench.err
echo "rc=$?"
cat gpurun_out/r2tp_bench.json | cut -c1-300
head -14 gpurun_out/r2tp_trace_rank.txt.rank0
timeout 600 python scripts/llm_bench.py --waves 3 2>/dev/null | cut -c1-300
