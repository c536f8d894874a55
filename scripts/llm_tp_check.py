"""torchrun worker (2 ranks, one GPU each): the tensor-parallel pair must reproduce the single-GPU run of the same
randomly initialised model.  The two layouts differ in where bf16 rounding happens (a prefill partial sum is
rounded per rank before the exchange), so the bar is the bf16 one of tests/test_gpu_llm.py: first-step logits within
2e-2 of the largest |logit|, the first sampled token equal wherever the top-2 margin is clear, and most greedy
continuations identical."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clearml_serving_b200 import llm as L, native  # noqa: E402


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("gloo")
    spec = L.LlamaSpec(vocab_size=2048, hidden_size=1024, intermediate_size=2048, num_hidden_layers=3,
                       num_attention_heads=8, num_key_value_heads=4, head_dim=128)
    rng = np.random.default_rng(5)
    prompts = [rng.integers(0, spec.vocab_size, n) for n in (9, 33, 64, 130)]
    n_new = 12
    def run(tp_size, tp_rank):
        eng = L.LlmEngine(spec, device=local, max_batch=4, max_ctx=256, tp_size=tp_size, tp_rank=tp_rank)
        eng.init_random(seed=3, std=0.05)
        eng.llm.keep_logits(True)
        eng.llm.prefill(prompts)
        first = eng.llm.logits()
        eng.llm.decode(n_new - 1)
        toks = eng.llm.tokens(n_new)
        if tp_size == 2:
            dist.barrier()          # the peer may still be reading this rank's exchange block
        eng.close()
        return first, toks

    lg2, toks2 = run(2, rank)
    both = [None, None]
    dist.all_gather_object(both, lg2)
    lg2 = np.concatenate(both, axis=1)
    lg1, toks1 = run(1, 0)          # every rank also runs the whole model on its own GPU
    scale = np.abs(lg1).max()
    err = np.abs(lg1 - lg2).max()
    top2 = np.sort(lg1, axis=1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > 4e-2 * scale
    first_ok = bool((toks1[clear, 0] == toks2[clear, 0]).all())
    same = (toks1 == toks2).all(axis=1)
    print("rank", rank, "first-step logit err", err, "scale", scale, "clear margins", int(clear.sum()), "first tokens ok", first_ok,
          "identical continuations", int(same.sum()), "of", len(prompts), flush=True)
    ok = err <= 2e-2 * scale and first_ok and same.sum() * 2 >= len(prompts)
    flags = [None, None]
    dist.all_gather_object(flags, bool(ok))
    dist.barrier()
    if rank == 0 and all(flags):
        print("TP2 OK", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if all(flags) else 1)


if __name__ == "__main__":
    main()
