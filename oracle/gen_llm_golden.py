"""Generates tests/golden/llama_tiny.npz with HuggingFace transformers' LlamaForCausalLM (fp32, CPU) -- the
network vLLM runs for the reference's LLM endpoint (see oracle/llm_oracle.py header).  Run in the build
container: `python oracle/gen_llm_golden.py`; the GPU box only reads the committed .npz."""
import os
import sys

import numpy as np
import torch
from transformers import LlamaConfig, LlamaForCausalLM

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clearml_serving_b200.llm import LlamaSpec, random_state_dict  # noqa: E402

SPEC = LlamaSpec(vocab_size=1024, hidden_size=512, intermediate_size=1024, num_hidden_layers=2,
                 num_attention_heads=4, num_key_value_heads=2, head_dim=128, rope_theta=500000.0, rms_norm_eps=1e-5)
SEED, STD, N_NEW = 7, 0.05, 8
PROMPT_LENS = (5, 17, 70)


def main():
    sd = random_state_dict(SPEC, seed=SEED, std=STD)
    cfg = LlamaConfig(vocab_size=SPEC.vocab_size, hidden_size=SPEC.hidden_size, intermediate_size=SPEC.intermediate_size,
                      num_hidden_layers=SPEC.num_hidden_layers, num_attention_heads=SPEC.num_attention_heads,
                      num_key_value_heads=SPEC.num_key_value_heads, head_dim=SPEC.head_dim, rope_theta=SPEC.rope_theta,
                      rms_norm_eps=SPEC.rms_norm_eps, max_position_embeddings=2048, tie_word_embeddings=False,
                      attention_bias=False, mlp_bias=False, hidden_act="silu")
    model = LlamaForCausalLM(cfg).float().eval()
    missing = model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    assert not [k for k in missing.missing_keys if "rotary" not in k], missing
    rng = np.random.default_rng(1)
    out = {"prompt_lens": np.asarray(PROMPT_LENS), "n_new": np.asarray(N_NEW), "seed": np.asarray(SEED), "std": np.asarray(STD)}
    for i, n in enumerate(PROMPT_LENS):
        prompt = rng.integers(0, SPEC.vocab_size, n)
        ids = torch.from_numpy(prompt)[None]
        toks, logits = [], []
        with torch.no_grad():
            for _ in range(N_NEW):
                lg = model(input_ids=ids).logits[0, -1].numpy()
                nxt = int(lg.argmax())
                toks.append(nxt)
                logits.append(lg)
                ids = torch.cat([ids, torch.tensor([[nxt]])], dim=1)
        out["prompt_%d" % i] = prompt.astype(np.int32)
        out["tokens_%d" % i] = np.asarray(toks, np.int32)
        out["logits_%d" % i] = np.stack(logits).astype(np.float32)
    path = os.path.join(ROOT, "tests", "golden", "llama_tiny.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
