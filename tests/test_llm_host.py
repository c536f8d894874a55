"""CPU tests of the LLM endpoint's host side: the oracle against the HuggingFace golden, weight sharding,
bf16 conversion."""
import os

import numpy as np

from clearml_serving_b200 import llm as L
from oracle import llm_oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden", "llama_tiny.npz")
SPEC = L.LlamaSpec(vocab_size=1024, hidden_size=512, intermediate_size=1024, num_hidden_layers=2,
                   num_attention_heads=4, num_key_value_heads=2, head_dim=128, rope_theta=500000.0, rms_norm_eps=1e-5)


def _gold():
    g = np.load(GOLD)
    return g, L.random_state_dict(SPEC, seed=int(g["seed"]), std=float(g["std"]))


def test_oracle_matches_transformers_golden():
    g, sd = _gold()
    for i in range(len(g["prompt_lens"])):
        toks, logits = llm_oracle.greedy_generate(sd, SPEC, g["prompt_%d" % i], int(g["n_new"]))
        assert np.array_equal(toks, g["tokens_%d" % i])
        ref = g["logits_%d" % i]
        assert np.abs(logits - ref).max() <= 2e-4 * np.abs(ref).max()


def test_bf16_round_trip_and_rounding():
    x = np.array([1.0, -2.5, 3.14159, 1e-8, 65504.0, 1.00390625, 1.01171875], np.float32)
    b = L.to_bf16_bits(x)
    back = L.from_bf16_bits(b)
    assert np.all(np.abs(back - x) <= np.abs(x) * 2.0 ** -8)
    assert np.array_equal(L.to_bf16_bits(back), b)                       # idempotent
    assert L.from_bf16_bits(L.to_bf16_bits(np.float32([1.00390625])))[0] == 1.0        # tie -> even (down)
    assert L.from_bf16_bits(L.to_bf16_bits(np.float32([1.01171875])))[0] == 1.015625   # tie -> even (up)


def test_init_value_statistics_and_block_consistency():
    full = L.init_value(3, 17, 64, 96, std=0.02)
    assert abs(float(full.mean())) < 2e-3 and abs(float(full.std()) - 0.02) < 2e-3
    blk = L.init_value(3, 17, 16, 32, row0=8, col0=40, std=0.02)
    assert np.array_equal(blk, full[8:24, 40:72])


def test_tensor_parallel_shards_reassemble():
    _, sd = _gold()
    s1 = L.shard_state_dict(sd, SPEC, 1, 0)
    a, b = L.shard_state_dict(sd, SPEC, 2, 0), L.shard_state_dict(sd, SPEC, 2, 1)
    hd, hq, hk, ir = 128, 2, 1, 512
    for l in range(SPEC.num_hidden_layers):
        q = np.concatenate([a[("wqkv", l)][:hq * hd], b[("wqkv", l)][:hq * hd]])
        k = np.concatenate([a[("wqkv", l)][hq * hd:(hq + hk) * hd], b[("wqkv", l)][hq * hd:(hq + hk) * hd]])
        v = np.concatenate([a[("wqkv", l)][(hq + hk) * hd:], b[("wqkv", l)][(hq + hk) * hd:]])
        assert np.array_equal(np.concatenate([q, k, v]), s1[("wqkv", l)])
        assert np.array_equal(np.concatenate([a[("wo", l)], b[("wo", l)]], axis=1), s1[("wo", l)])
        (ga, ua), (gb, ub) = L.split_gate_up(a[("wgu", l)]), L.split_gate_up(b[("wgu", l)])
        g1, u1 = L.split_gate_up(s1[("wgu", l)])
        assert np.array_equal(np.concatenate([ga, gb]), g1) and np.array_equal(np.concatenate([ua, ub]), u1)
        assert np.array_equal(g1, L.to_bf16_bits(sd["model.layers.%d.mlp.gate_proj.weight" % l]))
        assert np.array_equal(L.interleave_gate_up(g1, u1), s1[("wgu", l)])
        assert np.array_equal(s1[("wgu", l)][32:64], L.to_bf16_bits(sd["model.layers.%d.mlp.up_proj.weight" % l][:32]))
        assert np.array_equal(np.concatenate([a[("wdown", l)], b[("wdown", l)]], axis=1), s1[("wdown", l)])
    assert np.array_equal(np.concatenate([a[("lm_head", 0)], b[("lm_head", 0)]]), s1[("lm_head", 0)])
    assert np.array_equal(a[("embed", 0)], s1[("embed", 0)])


def test_row_parallel_partials_sum_to_full_projection():
    # the algebra the fused all-reduce relies on: x @ Wo^T == sum over ranks of x[:, cols_r] @ Wo[:, cols_r]^T
    _, sd = _gold()
    w = sd["model.layers.0.self_attn.o_proj.weight"]
    x = np.random.default_rng(0).standard_normal((5, w.shape[1])).astype(np.float32)
    half = w.shape[1] // 2
    parts = x[:, :half] @ w[:, :half].T + x[:, half:] @ w[:, half:].T
    assert np.allclose(parts, x @ w.T, rtol=1e-5, atol=1e-5)


def test_spec_counts_llama3_8b():
    s = L.LlamaSpec.llama3_8b()
    assert s.n_params() == 8030261248          # meta-llama/Meta-Llama-3-8B parameter count
    assert abs(s.flops_per_token() - 2 * 7.505e9) < 2e7
