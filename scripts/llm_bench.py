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


def trace_steps(m, prompts, out_path, steps=4):
    """Kernel timeline of graph-launched decode steps (CUPTI through torch.profiler; nothing under it is a bench number).
    For every kernel: start, end, and how far it moved the completion frontier (end - max(previous ends)): the frontier
    increments of a step add up to the step time, so they attribute it to kernels INCLUDING launch gaps and dependency stalls."""
    import collections
    import tempfile
    import torch
    from torch.profiler import ProfilerActivity, profile
    m.prefill(prompts)
    m.decode(4, use_graph=True)
    m.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        m.decode(steps, use_graph=True)
        m.synchronize()
        torch.cuda.synchronize()
    tmp = tempfile.mktemp(suffix=".json")
    prof.export_chrome_trace(tmp)
    ev = [e for e in json.load(open(tmp))["traceEvents"] if e.get("cat") == "kernel"]
    ev.sort(key=lambda e: e["ts"])
    names = [e["name"].split("(")[0].split("<")[0].replace("b2s::", "") for e in ev]
    ends = [i for i, n in enumerate(names) if "step_end" in n]
    lines = []
    if len(ends) >= 3:
        lo, hi = ends[-3] + 1, ends[-2] + 1          # the second-to-last traced step
        seq = ev[lo:hi]
        t0 = ev[lo - 1]["ts"] + ev[lo - 1]["dur"]
        frontier = t0
        agg = collections.OrderedDict()
        per = []
        idx_in_layer = collections.Counter()
        for e, n in zip(seq, names[lo:hi]):
            st, en = e["ts"], e["ts"] + e["dur"]
            inc = max(0.0, en - frontier)
            key = n + "#" + str(idx_in_layer[n] % (4 if "skinny" in n else 2 if "reduce_rms" in n else 1)) if ("skinny" in n or "reduce_rms" in n) else n
            idx_in_layer[n] += 1
            a_ = agg.setdefault(key, [0, 0.0, 0.0, 0.0])
            a_[0] += 1; a_[1] += inc; a_[2] += e["dur"]; a_[3] += st - frontier
            per.append((n, st - t0, en - t0, e["dur"], inc, st - frontier))
            frontier = max(frontier, en)
        total = frontier - t0
        lines.append("step time (frontier) %.1f us over %d kernels" % (total, len(seq)))
        lines.append("%-44s %5s %12s %12s %14s" % ("kernel#slot", "n", "frontier us", "sum dur us", "mean start-frontier"))
        for k, v in agg.items():
            lines.append("%-44s %5d %12.1f %12.1f %14.2f" % (k, v[0], v[1], v[2], v[3] / v[0]))
        lines.append("")
        lines.append("first 3 layers, per kernel: name, start, end, dur, frontier increment, start - frontier (negative = overlapped its predecessor)")
        for r in per[:27]:
            lines.append("%-40s %9.2f %9.2f %8.2f %8.2f %8.2f" % r)
    else:
        lines.append("trace holds %d kernels, %d step ends" % (len(ev), len(ends)))
    open(out_path, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]), file=sys.stderr)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--prompt", type=int, default=512)
    ap.add_argument("--gen", type=int, default=128)
    ap.add_argument("--waves", type=int, default=3)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--timing", action="store_true", help="per-kernel device times of eager decode steps (stderr)")
    ap.add_argument("--trace", default="", help="CUPTI (torch.profiler) kernel timeline of graph-mode decode steps -> this text file")
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
    if a.trace:
        trace_steps(m, prompts, a.trace if world == 1 else "{}.rank{}".format(a.trace, rank))
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
