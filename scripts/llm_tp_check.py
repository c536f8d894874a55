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


def check(tp_rank, local, group=None):
    """Collective over the 2 ranks of `group` (None: the world).  Returns (ok, figures) on every rank."""
    spec = L.LlamaSpec(vocab_size=2048, hidden_size=1024, intermediate_size=2048, num_hidden_layers=3,
                       num_attention_heads=8, num_key_value_heads=4, head_dim=128)
    rng = np.random.default_rng(5)
    prompts = [rng.integers(0, spec.vocab_size, n) for n in (9, 33, 64, 130)]
    n_new = 12

    def run(tp_size, rank_in_pair):
        eng = L.LlmEngine(spec, device=local, max_batch=4, max_ctx=256, tp_size=tp_size, tp_rank=rank_in_pair, tp_group=group)
        eng.init_random(seed=3, std=0.05)
        eng.llm.keep_logits(True)
        eng.llm.prefill(prompts)
        first = eng.llm.logits()
        eng.llm.decode(n_new - 1)
        toks = eng.llm.tokens(n_new)
        if tp_size == 2:
            dist.barrier(group=group)          # the peer may still be reading this rank's exchange block
        eng.close()
        return first, toks

    lg2, toks2 = run(2, tp_rank)
    both = [None, None]
    dist.all_gather_object(both, (lg2, toks2), group=group)
    pair_tokens_equal = bool(np.array_equal(both[0][1], both[1][1]))
    lg2 = np.concatenate([both[0][0], both[1][0]], axis=1)
    lg1, toks1 = run(1, 0)          # every rank also runs the whole model on its own GPU
    scale = float(np.abs(lg1).max())
    err = float(np.abs(lg1 - lg2).max())
    top2 = np.sort(lg1, axis=1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > 4e-2 * scale
    first_ok = bool((toks1[clear, 0] == toks2[clear, 0]).all())
    same = (toks1 == toks2).all(axis=1)
    ok = err <= 2e-2 * scale and first_ok and same.sum() * 2 >= len(prompts) and pair_tokens_equal
    flags = [None, None]
    dist.all_gather_object(flags, bool(ok), group=group)
    dist.barrier(group=group)
    return all(flags), dict(first_step_logit_err=err, logit_scale=scale, bar=2e-2 * scale, clear_margins=int(clear.sum()),
                            first_tokens_ok=first_ok, identical_continuations=int(same.sum()), prompts=len(prompts),
                            pair_ranks_same_tokens=pair_tokens_equal)


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("gloo")
    ok, fig = check(rank, local)
    print("rank", rank, fig, flush=True)
    if rank == 0 and ok:
        print("TP2 OK", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
