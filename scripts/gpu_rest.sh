#!/bin/bash
timeout 150 python bench.py --no-bert --no-resnet --no-llama --cpu-seconds 1 2>gpurun_out/bench_rest.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); p=d['plugin']; print('closed', round(p['closed_loop_req_s'])); print('rest', json.dumps(p.get('rest'))[:900]); print('clocks', d['clocks'])"; tail -3 gpurun_out/bench_rest.err
