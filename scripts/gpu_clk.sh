#!/bin/bash
timeout 200 python bench.py --no-bert --no-resnet --no-llama --no-plugin --cpu-seconds 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['clocks'], d['gpu_launches'])"
