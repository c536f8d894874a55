#!/bin/bash
cd /root/repo
echo "== stream"; timeout 300 python scripts/llm_invariance_probe.py 2>&1 | tail -4
echo "== per-sequence CTAs"; B2S_LLM_ATTN_STREAM=0 timeout 300 python scripts/llm_invariance_probe.py 2>&1 | tail -4
