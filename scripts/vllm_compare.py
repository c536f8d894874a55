"""The comparator SURVEY.md 2.1 names for BASELINE configs[4]: vLLM (the engine the reference wraps,
clearml_serving/serving/preprocess_service.py:632-683,1097-1348) on the SAME box and the SAME workload -- Llama-3-8B
architecture, random ("dummy") bf16 weights, 32 prompts x 512 tokens, 128 new tokens each, greedy, ignore_eos.

    python scripts/vllm_compare.py [--tp 1|2] [--waves 3] [--out gpurun_out/vllm_compare.json]

Prints one JSON line: requests/s and generated tokens/s of a closed 32-request wave (the best of `--waves`), plus
mean TTFT of the wave when vLLM reports per-request metrics.  LIBRARY code: nothing here is part of the product path."""
import argparse
import json
import os
import sys
import tempfile
import time


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tp", type=int, default=1)
    ap.add_argument("--waves", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--prompt-len", type=int, default=512)
    ap.add_argument("--gen", type=int, default=128)
    ap.add_argument("--out", default="")
    ap.add_argument("--eager", action="store_true")
    args = ap.parse_args()
    os.environ.setdefault("HF_HUB_OFFLINE", "1")
    os.environ.setdefault("TRANSFORMERS_OFFLINE", "1")
    os.environ.setdefault("VLLM_NO_USAGE_STATS", "1")
    os.environ.setdefault("TOKENIZERS_PARALLELISM", "false")
    import numpy as np
    from transformers import LlamaConfig
    d = tempfile.mkdtemp(prefix="llama3_8b_cfg_")
    LlamaConfig(vocab_size=128256, hidden_size=4096, intermediate_size=14336, num_hidden_layers=32, num_attention_heads=32,
                num_key_value_heads=8, max_position_embeddings=8192, rope_theta=500000.0, rms_norm_eps=1e-5,
                torch_dtype="bfloat16", architectures=["LlamaForCausalLM"]).save_pretrained(d)
    t0 = time.time()
    import vllm
    from vllm import LLM, SamplingParams
    llm = LLM(model=d, load_format="dummy", dtype="bfloat16", tensor_parallel_size=args.tp, max_model_len=args.prompt_len + args.gen + 16,
              max_num_seqs=args.batch, gpu_memory_utilization=0.6, skip_tokenizer_init=True, enforce_eager=args.eager, seed=0)
    t_load = time.time() - t0
    rng = np.random.default_rng(1)
    prompts = [{"prompt_token_ids": rng.integers(0, 128256, args.prompt_len).tolist()} for _ in range(args.batch)]
    sp = SamplingParams(max_tokens=args.gen, ignore_eos=True, temperature=0.0, detokenize=False)
    llm.generate(prompts[:4], SamplingParams(max_tokens=8, ignore_eos=True, temperature=0.0, detokenize=False), use_tqdm=False)   # warm-up
    waves = []
    ttft = []
    for _ in range(args.waves):
        t = time.perf_counter()
        outs = llm.generate(prompts, sp, use_tqdm=False)
        dt = time.perf_counter() - t
        waves.append(dt)
        assert all(len(o.outputs[0].token_ids) == args.gen for o in outs)
        for o in outs:
            m = getattr(o, "metrics", None)
            if m is not None and getattr(m, "first_token_time", None) and getattr(m, "arrival_time", None):
                ttft.append(m.first_token_time - m.arrival_time)
    best = min(waves)
    res = dict(engine="vllm " + vllm.__version__, tensor_parallel=args.tp, batch=args.batch, prompt_len=args.prompt_len, gen=args.gen,
               requests_per_s=args.batch / best, gen_tokens_per_s=args.batch * args.gen / best, wave_s=[round(w, 4) for w in waves],
               mean_ttft_s=(sum(ttft) / len(ttft)) if ttft else None, load_s=round(t_load, 1), enforce_eager=bool(args.eager),
               weights="dummy (random) bf16, Llama-3-8B architecture")
    line = json.dumps(res)
    print(line)
    if args.out:
        with open(args.out, "w") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
