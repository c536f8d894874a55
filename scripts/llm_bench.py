"""Llama-3-8B (random init, bf16) wave benchmark: 32 prompts x 512 tokens, 128 new tokens (BASELINE.json configs[4]).
Single process = TP 1; under torchrun with 2 ranks = one tensor-parallel pair.  Device-timed with CUDA events."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clearml_serving_b200 import llm as L  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--prompt", type=int, default=512)
    ap.add_argument("--gen", type=int, default=128)
    ap.add_argument("--waves", type=int, default=3)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--timing", action="store_true", help="per-kernel device times of eager decode steps (stderr)")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")
    spec = L.LlamaSpec.llama3_8b()
    spec.num_hidden_layers = a.layers
    t0 = time.time()
    eng = L.LlmEngine(spec, device=local, max_batch=a.batch, max_ctx=a.prompt + a.gen + 16, max_tokens=a.batch * a.prompt,
                      tp_size=2 if world == 2 else 1, tp_rank=rank if world == 2 else 0)
    eng.init_random(seed=0, std=0.02)
    t_init = time.time() - t0
    rng = np.random.default_rng(1)
    prompts = [rng.integers(0, spec.vocab_size, a.prompt) for _ in range(a.batch)]
    m = eng.llm
    res = []
    for w in range(a.waves):
        m.flush_l2()
        m.record(0)
        m.prefill(prompts)
        m.record(1)
        m.decode(a.gen - 1, use_graph=not a.no_graph)
        m.record(2)
        toks = m.tokens(a.gen)
        res.append((m.elapsed_ms(0, 1), m.elapsed_ms(1, 2)))
    if a.timing:
        m.prefill(prompts)
        m.decode(12, use_graph=2)
        m.synchronize()
    pre, dec = np.median([r[0] for r in res[1:] or res]), np.median([r[1] for r in res[1:] or res])
    tp = 2 if world == 2 else 1
    step_ms = dec / (a.gen - 1)
    wbytes = (spec.n_params() - spec.vocab_size * spec.hidden_size) * 2 / tp   # streamed per decode step (embedding is gathered)
    kv_bytes = a.batch * (a.prompt + a.gen / 2) * spec.num_hidden_layers * 2 * spec.num_key_value_heads * 128 * 2 / tp
    pre_flops = a.batch * a.prompt * spec.flops_per_token() / tp
    if rank == 0:
        print(json.dumps({
            "workload": "llama3-8b-random bf16 tp{} batch {} prompt {} gen {} layers {}".format(tp, a.batch, a.prompt, a.gen, a.layers),
            "init_s": round(t_init, 1), "prefill_ms": round(float(pre), 2), "decode_ms": round(float(dec), 2),
            "decode_step_ms": round(float(step_ms), 4),
            "req_per_s": round(a.batch / ((pre + dec) / 1e3), 2), "gen_tok_per_s": round(a.batch * a.gen / ((pre + dec) / 1e3), 1),
            "prefill_tflops": round(pre_flops / (pre / 1e3) / 1e12, 1),
            "decode_hbm_gbps": round((wbytes + kv_bytes) / (step_ms / 1e3) / 1e9, 1),
            "waves_ms": [[round(x, 2) for x in r] for r in res], "tokens_head": toks[0, :8].tolist()}), flush=True)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
