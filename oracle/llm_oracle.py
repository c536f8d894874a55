"""CPU oracle for the LLM endpoint (BASELINE.json configs[4]) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this file;
the product path (clearml_serving_b200/) never does.

What it restates: the Llama decoder forward pass the reference's vLLM engine computes for this endpoint
(reference call sites: clearml_serving/serving/preprocess_service.py:1097-1348 builds the vLLM engine from the
endpoint's auxiliary_cfg, examples/vllm/preprocess.py; the arithmetic itself lives in third-party code --
vllm 0.x `vllm/model_executor/models/llama.py`, numerically the same network as HuggingFace transformers
`models/llama/modeling_llama.py`: LlamaRMSNorm, rotate_half RoPE, grouped-query causal attention, SwiGLU MLP).
Plain fp32 numpy, one full forward per generated token (no KV cache), so it only suits tiny configurations.

Pinned: tests/golden/llama_tiny.npz holds logits and greedy tokens produced by transformers'
LlamaForCausalLM (fp32, CPU) in this container with oracle/gen_llm_golden.py; tests/test_llm_host.py checks
this restatement against them.
"""
import numpy as np


def _rms(x, w, eps):
    # modeling_llama.LlamaRMSNorm.forward: x * rsqrt(mean(x^2) + eps) * weight
    var = np.mean(x.astype(np.float32) ** 2, axis=-1, keepdims=True)
    return x * (1.0 / np.sqrt(var + np.float32(eps))) * w


def _rope_tables(n_pos, head_dim, theta):
    # modeling_llama.LlamaRotaryEmbedding: inv_freq = 1 / theta^(2i/d); emb = cat(freqs, freqs)
    inv_freq = (1.0 / (np.float32(theta) ** (np.arange(0, head_dim, 2, dtype=np.float32) / np.float32(head_dim)))).astype(np.float32)
    ang = np.arange(n_pos, dtype=np.float32)[:, None] * inv_freq[None, :]
    return np.cos(ang).astype(np.float32), np.sin(ang).astype(np.float32)


def _apply_rope(x, cos, sin):
    # x: [T, heads, d]; rotate_half convention: out = x * cos + cat(-x2, x1) * sin with pairs (i, i + d/2)
    d = x.shape[-1] // 2
    x1, x2 = x[..., :d], x[..., d:]
    c, s = cos[:, None, :], sin[:, None, :]
    return np.concatenate([x1 * c - x2 * s, x2 * c + x1 * s], axis=-1)


def forward_logits(sd, spec, tokens):
    """tokens: 1-D int array (one sequence) -> fp32 logits [T, vocab]"""
    T = len(tokens)
    hd, hq, hk = spec.head_dim, spec.num_attention_heads, spec.num_key_value_heads
    g = hq // hk
    x = sd["model.embed_tokens.weight"][np.asarray(tokens)].astype(np.float32)
    cos, sin = _rope_tables(T, hd, spec.rope_theta)
    mask = np.triu(np.full((T, T), -np.inf, dtype=np.float32), k=1)
    for l in range(spec.num_hidden_layers):
        p = "model.layers.{}.".format(l)
        xn = _rms(x, sd[p + "input_layernorm.weight"], spec.rms_norm_eps)
        q = (xn @ sd[p + "self_attn.q_proj.weight"].T).reshape(T, hq, hd)
        k = (xn @ sd[p + "self_attn.k_proj.weight"].T).reshape(T, hk, hd)
        v = (xn @ sd[p + "self_attn.v_proj.weight"].T).reshape(T, hk, hd)
        q, k = _apply_rope(q, cos, sin), _apply_rope(k, cos, sin)
        ctx = np.empty((T, hq, hd), dtype=np.float32)
        for h in range(hq):
            s = (q[:, h] @ k[:, h // g].T) / np.float32(np.sqrt(hd)) + mask
            s = s - s.max(axis=-1, keepdims=True)
            pr = np.exp(s)
            pr = pr / pr.sum(axis=-1, keepdims=True)
            ctx[:, h] = pr @ v[:, h // g]
        x = x + ctx.reshape(T, hq * hd) @ sd[p + "self_attn.o_proj.weight"].T
        xn = _rms(x, sd[p + "post_attention_layernorm.weight"], spec.rms_norm_eps)
        gate = xn @ sd[p + "mlp.gate_proj.weight"].T
        up = xn @ sd[p + "mlp.up_proj.weight"].T
        x = x + ((gate / (1.0 + np.exp(-gate))) * up) @ sd[p + "mlp.down_proj.weight"].T
    x = _rms(x, sd["model.norm.weight"], spec.rms_norm_eps)
    head = sd["lm_head.weight"] if "lm_head.weight" in sd else sd["model.embed_tokens.weight"]
    return (x @ head.T).astype(np.float32)


def greedy_generate(sd, spec, prompt, n_new):
    """-> (tokens [n_new], logits of each sampling step [n_new, vocab])"""
    toks = list(int(t) for t in prompt)
    out, lg = [], []
    for _ in range(n_new):
        logits = forward_logits(sd, spec, np.asarray(toks))[-1]
        nxt = int(np.argmax(logits))
        out.append(nxt)
        lg.append(logits)
        toks.append(nxt)
    return np.asarray(out, dtype=np.int32), np.stack(lg)
