"""torchrun worker (2 ranks, one GPU each): the tensor-parallel pair must reproduce the single-GPU run of the same
randomly initialised model -- logits within 1e-3 of scale (only the fp32 summation order differs), same tokens."""
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
    # TP=2 pair
    eng = L.LlmEngine(spec, device=local, max_batch=4, max_ctx=256, tp_size=2, tp_rank=rank)
    eng.init_random(seed=3, std=0.05)
    eng.llm.keep_logits(True)
    toks2 = eng.generate(prompts, n_new)
    lg2 = eng.llm.logits()
    both = [None, None]
    dist.all_gather_object(both, lg2)
    lg2 = np.concatenate(both, axis=1)
    eng.close()
    # TP=1 on each rank's own GPU
    ref = L.LlmEngine(spec, device=local, max_batch=4, max_ctx=256)
    ref.init_random(seed=3, std=0.05)
    ref.llm.keep_logits(True)
    toks1 = ref.generate(prompts, n_new)
    lg1 = ref.llm.logits()
    ref.close()
    same = (toks1 == toks2).all(axis=1)
    scale = np.abs(lg1).max()
    err = np.abs(lg1 - lg2)[same].max() if same.any() else float("inf")
    print("rank", rank, "identical sequences", int(same.sum()), "of", len(prompts), "logit err", err, "scale", scale, flush=True)
    ok = same.sum() >= len(prompts) - 1 and err <= 1e-3 * scale
    flags = [None, None]
    dist.all_gather_object(flags, bool(ok))
    dist.barrier()
    if rank == 0 and all(flags):
        print("TP2 OK", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if all(flags) else 1)


if __name__ == "__main__":
    main()
