"""Developer probe: is a sequence's decode output bit-identical whatever shares the batch with it?  (tiny model, logits per step)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clearml_serving_b200 import llm as L, native  # noqa: E402

SPEC = L.LlamaSpec(vocab_size=1024, hidden_size=512, intermediate_size=1024, num_hidden_layers=2,
                   num_attention_heads=4, num_key_value_heads=2, head_dim=128, rope_theta=500000.0, rms_norm_eps=1e-5)


def run(prompts, row, steps, graph):
    m = native.Llm(device=0, vocab=SPEC.vocab_size, hidden=SPEC.hidden_size, inter=SPEC.intermediate_size,
                   n_layers=SPEC.num_hidden_layers, n_heads=SPEC.num_attention_heads, n_kv_heads=SPEC.num_key_value_heads,
                   max_batch=4, max_ctx=256, rope_theta=SPEC.rope_theta, rms_eps=SPEC.rms_norm_eps)
    try:
        m.init_random(5, 0.05)
        m.keep_logits(True)
        m.prefill(prompts)
        out = [m.logits()[row].copy()]
        for s in range(steps):
            m.decode(1, use_graph=graph)
            out.append(m.logits()[row].copy())
        return out
    finally:
        m.free()


def main():
    rng = np.random.default_rng(17)
    p3 = rng.integers(0, SPEC.vocab_size, 63)
    others = [rng.integers(0, SPEC.vocab_size, n) for n in (5, 70, 130)]
    for graph in (False, True):
        a = run([p3], 0, 70, graph)
        b = run([others[0], others[1], p3, others[2]], 2, 70, graph)
        c = run([p3], 0, 70, graph)
        first_ab = next((i for i, (x, y) in enumerate(zip(a, b)) if not np.array_equal(x, y)), None)
        first_ac = next((i for i, (x, y) in enumerate(zip(a, c)) if not np.array_equal(x, y)), None)
        print("graph", graph, "solo vs batch: first differing step", first_ab,
              "max diff there", None if first_ab is None else float(np.abs(a[first_ab] - b[first_ab]).max()),
              "| solo vs solo:", first_ac, flush=True)


if __name__ == "__main__":
    main()
