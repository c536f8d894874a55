"""Host side of the LLM endpoint (BASELINE.json configs[4]: Llama-3-8B bf16, tensor-parallel 2, max_batch 32).

The reference serves this endpoint by wrapping vLLM (`VllmPreprocessRequest`,
clearml_serving/serving/preprocess_service.py:1097-1348; engine args such as `tensor_parallel_size` come from
the endpoint's `auxiliary_cfg`, examples/vllm/preprocess.py).  Here the model runs on libb200serve.so
(csrc/llm.cu); this module is the Python around it:

  * `LlamaSpec`            -- the architecture numbers (HF `LlamaConfig` names)
  * `shard_state_dict`     -- HF-named fp32/bf16 weights -> this rank's fused, Megatron-split bf16 tensors
  * `init_value`           -- numpy twin of the on-device deterministic initialiser (bit-exact)
  * `attach_tensor_parallel` -- swaps the cudaIpc handles of a rank pair over torch.distributed
  * `LlmEngine`            -- prompt waves of <= max_batch sequences: prefill + CUDA-graph decode, greedy

One process drives one GPU; a tensor-parallel pair is two processes issuing the same calls (rank 0 broadcasts
the prompts).  There is no CPU fallback: every method raises if the CUDA library is missing.
"""
from dataclasses import dataclass

import numpy as np

from . import native


@dataclass
class LlamaSpec:
    vocab_size: int = 128256
    hidden_size: int = 4096
    intermediate_size: int = 14336
    num_hidden_layers: int = 32
    num_attention_heads: int = 32
    num_key_value_heads: int = 8
    head_dim: int = 128
    rope_theta: float = 500000.0
    rms_norm_eps: float = 1e-5

    @classmethod
    def llama3_8b(cls):
        return cls()

    @classmethod
    def from_hf_config(cls, c):
        hd = getattr(c, "head_dim", None) or c.hidden_size // c.num_attention_heads
        theta = getattr(c, "rope_theta", None)
        if theta is None:
            theta = (getattr(c, "rope_parameters", None) or {}).get("rope_theta", 10000.0)
        return cls(c.vocab_size, c.hidden_size, c.intermediate_size, c.num_hidden_layers, c.num_attention_heads,
                   c.num_key_value_heads, hd, float(theta), float(c.rms_norm_eps))

    def n_params(self):
        h, i, v = self.hidden_size, self.intermediate_size, self.vocab_size
        qkv = (self.num_attention_heads + 2 * self.num_key_value_heads) * self.head_dim * h
        per_layer = qkv + self.num_attention_heads * self.head_dim * h + 3 * i * h + 2 * h
        return 2 * v * h + self.num_hidden_layers * per_layer + h

    def flops_per_token(self):
        """dense matmul FLOPs of one token through the stack + lm_head (no attention term)"""
        h, i = self.hidden_size, self.intermediate_size
        qkv = (self.num_attention_heads + 2 * self.num_key_value_heads) * self.head_dim * h
        per_layer = qkv + self.num_attention_heads * self.head_dim * h + 3 * i * h
        return 2 * (self.num_hidden_layers * per_layer + self.vocab_size * h)


def to_bf16_bits(a):
    """float32 -> bf16 bit patterns (uint16), round to nearest even (what __float2bfloat16_rn does)"""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    rounded = u + (np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1)))
    return (rounded >> np.uint32(16)).astype(np.uint16)


def from_bf16_bits(b):
    return (np.ascontiguousarray(b, dtype=np.uint16).astype(np.uint32) << np.uint32(16)).view(np.float32)


def _mix64(z):
    z = z + np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def init_value(seed, tensor_id, rows, cols, row0=0, col0=0, std=0.02):
    """numpy twin of llm_init_value (csrc/llm.cu): fp32 values of the [rows, cols] block whose top-left global
    coordinate is (row0, col0) of tensor `tensor_id`"""
    with np.errstate(over="ignore"):
        r = (np.arange(rows, dtype=np.uint64) + np.uint64(row0))[:, None]
        c = (np.arange(cols, dtype=np.uint64) + np.uint64(col0))[None, :]
        key = np.uint64(seed) ^ (np.uint64(tensor_id) << np.uint64(48)) ^ (r << np.uint64(24)) ^ c
        z = _mix64(key)
    m = np.uint64(0xFFFF)
    s = ((z & m) + ((z >> np.uint64(16)) & m) + ((z >> np.uint64(32)) & m) + ((z >> np.uint64(48)) & m)).astype(np.int64)
    s = (s - 2 * 65535).astype(np.float32)
    return s * np.float32(np.float32(std) * np.float32(np.float32(1.7320508) / np.float32(65536.0)))


TENSOR_IDS = {"embed": 0, "lm_head": 1, "q": 0, "k": 1, "v": 2, "o": 3, "gate": 4, "up": 5, "down": 6}


def random_state_dict(spec, seed=0, std=0.02):
    """the full (unsharded) model `b2s_llm_init_random` creates, as HF-named float32 arrays (bf16-representable)"""
    h, hd = spec.hidden_size, spec.head_dim
    rb = lambda a: from_bf16_bits(to_bf16_bits(a))
    sd = {"model.embed_tokens.weight": rb(init_value(seed, 0, spec.vocab_size, h, std=std)),
          "lm_head.weight": rb(init_value(seed, 1, spec.vocab_size, h, std=std)),
          "model.norm.weight": np.ones(h, np.float32)}
    for l in range(spec.num_hidden_layers):
        tid = 16 + 8 * l
        p = "model.layers.{}.".format(l)
        sd[p + "self_attn.q_proj.weight"] = rb(init_value(seed, tid + 0, spec.num_attention_heads * hd, h, std=std))
        sd[p + "self_attn.k_proj.weight"] = rb(init_value(seed, tid + 1, spec.num_key_value_heads * hd, h, std=std))
        sd[p + "self_attn.v_proj.weight"] = rb(init_value(seed, tid + 2, spec.num_key_value_heads * hd, h, std=std))
        sd[p + "self_attn.o_proj.weight"] = rb(init_value(seed, tid + 3, h, spec.num_attention_heads * hd, std=std))
        sd[p + "mlp.gate_proj.weight"] = rb(init_value(seed, tid + 4, spec.intermediate_size, h, std=std))
        sd[p + "mlp.up_proj.weight"] = rb(init_value(seed, tid + 5, spec.intermediate_size, h, std=std))
        sd[p + "mlp.down_proj.weight"] = rb(init_value(seed, tid + 6, h, spec.intermediate_size, std=std))
        sd[p + "input_layernorm.weight"] = np.ones(h, np.float32)
        sd[p + "post_attention_layernorm.weight"] = np.ones(h, np.float32)
    return sd


def _np(x):
    if hasattr(x, "detach"):
        x = x.detach().float().cpu().numpy()
    return np.asarray(x, dtype=np.float32)


def interleave_gate_up(g, u):
    """[I, H] gate and up -> fused [2I, H] with rows interleaved in blocks of 32 (gate block, up block, ...): two
    consecutive 32-column chunks of the fused projection are (gate, up) of the same 32 outputs, which is what lets
    the GEMM epilogue apply SwiGLU before anything is written (csrc/gemm.cu epilogue_swiglu32)"""
    i, h = g.shape
    assert i % 32 == 0 and u.shape == g.shape
    return np.stack([g.reshape(i // 32, 32, h), u.reshape(i // 32, 32, h)], axis=1).reshape(2 * i, h)


def split_gate_up(fused):
    i2, h = fused.shape
    blocks = fused.reshape(i2 // 64, 2, 32, h)
    return blocks[:, 0].reshape(i2 // 2, h), blocks[:, 1].reshape(i2 // 2, h)


def shard_state_dict(sd, spec, tp_size=1, tp_rank=0):
    """HF-named weights -> {(name, layer): array} for this rank, in the layouts b2s_llm_tensor documents:
    QKV / gate / up split by OUTPUT rows (whole heads; gate and up fused and interleaved, see interleave_gate_up),
    O / down by INPUT columns, lm_head by vocabulary rows;
    bf16 tensors as uint16 bit patterns, norm weights float32."""
    hd = spec.head_dim
    hq, hk = spec.num_attention_heads // tp_size, spec.num_key_value_heads // tp_size
    ir, vr = spec.intermediate_size // tp_size, spec.vocab_size // tp_size
    r = tp_rank
    out = {("embed", 0): to_bf16_bits(_np(sd["model.embed_tokens.weight"])),
           ("final_norm", 0): _np(sd["model.norm.weight"])}
    head = sd["lm_head.weight"] if "lm_head.weight" in sd else sd["model.embed_tokens.weight"]   # tied embeddings
    out[("lm_head", 0)] = to_bf16_bits(_np(head)[r * vr:(r + 1) * vr])
    for l in range(spec.num_hidden_layers):
        p = "model.layers.{}.".format(l)
        q = _np(sd[p + "self_attn.q_proj.weight"])[r * hq * hd:(r + 1) * hq * hd]
        k = _np(sd[p + "self_attn.k_proj.weight"])[r * hk * hd:(r + 1) * hk * hd]
        v = _np(sd[p + "self_attn.v_proj.weight"])[r * hk * hd:(r + 1) * hk * hd]
        out[("wqkv", l)] = to_bf16_bits(np.concatenate([q, k, v], axis=0))
        out[("wo", l)] = to_bf16_bits(_np(sd[p + "self_attn.o_proj.weight"])[:, r * hq * hd:(r + 1) * hq * hd])
        g = _np(sd[p + "mlp.gate_proj.weight"])[r * ir:(r + 1) * ir]
        u = _np(sd[p + "mlp.up_proj.weight"])[r * ir:(r + 1) * ir]
        out[("wgu", l)] = to_bf16_bits(interleave_gate_up(g, u))
        out[("wdown", l)] = to_bf16_bits(_np(sd[p + "mlp.down_proj.weight"])[:, r * ir:(r + 1) * ir])
        out[("ln1", l)] = _np(sd[p + "input_layernorm.weight"])
        out[("ln2", l)] = _np(sd[p + "post_attention_layernorm.weight"])
    return out


def attach_tensor_parallel(llm, group=None):
    """Swap the cudaIpc handles of the exchange blocks inside a 2-rank torch.distributed group (any backend)
    and attach the peer's.  Collective: both ranks of the pair must call it."""
    import torch.distributed as dist
    handles = [None, None]
    dist.all_gather_object(handles, llm.comm_export(), group=group)
    me = dist.get_rank(group=group)
    llm.comm_attach(handles[1 - me])
    dist.barrier(group=group)   # nobody starts a step before both sides are attached


class LlmEngine(object):
    """Greedy generation in waves of up to `max_batch` prompts (BASELINE.json configs[4] workload shape:
    prompt 512, 128 new tokens, 32 sequences).  `tp_group`: a 2-rank torch.distributed group, or None (TP=1)."""

    def __init__(self, spec, device=0, max_batch=32, max_ctx=1024, max_tokens=None, tp_size=1, tp_rank=0, tp_group=None,
                 kv_pages=0):
        self.spec = spec
        self.max_batch, self.max_ctx = int(max_batch), int(max_ctx)
        self.max_prefill_tokens = int(max_tokens) if max_tokens else self.max_batch * self.max_ctx
        self.tp_size, self.tp_rank, self.tp_group = int(tp_size), int(tp_rank), tp_group
        self.llm = native.Llm(device=device, vocab=spec.vocab_size, hidden=spec.hidden_size, inter=spec.intermediate_size,
                              n_layers=spec.num_hidden_layers, n_heads=spec.num_attention_heads,
                              n_kv_heads=spec.num_key_value_heads, head_dim=spec.head_dim, max_batch=max_batch,
                              max_ctx=max_ctx, max_tokens=max_tokens, tp_size=tp_size, tp_rank=tp_rank,
                              rope_theta=spec.rope_theta, rms_eps=spec.rms_norm_eps, kv_pages=kv_pages)
        if self.tp_size == 2:
            attach_tensor_parallel(self.llm, tp_group)

    def init_random(self, seed=0, std=0.02):
        self.llm.init_random(seed, std)

    def load_state_dict(self, sd):
        for (name, layer), arr in shard_state_dict(sd, self.spec, self.tp_size, self.tp_rank).items():
            self.llm.load_tensor(name, layer, arr)

    def _check(self, prompts, max_new_tokens):
        if not prompts:
            raise ValueError("generate: empty prompt list")
        for p in prompts:
            if len(p) < 1 or len(p) + max_new_tokens > self.max_ctx:
                raise ValueError("generate: prompt of {} tokens + {} new exceeds max_ctx {}".format(
                    len(p), max_new_tokens, self.max_ctx))

    def generate(self, prompts, max_new_tokens, use_graph=True, on_progress=None, chunk=8):
        """prompts: list of token-id sequences -> int32 array [len(prompts), max_new_tokens].
        `on_progress(first_prompt_index, tokens[n_wave, n_done])`, when given, is called after the prefill (one token per
        sequence) and after every `chunk` decode steps: the same kernels, the step loop just comes up for air (one
        stream synchronisation per chunk) so that a serving front end can stream tokens while the wave runs."""
        max_new_tokens = int(max_new_tokens)
        self._check(prompts, max_new_tokens)
        out = np.empty((len(prompts), max_new_tokens), dtype=np.int32)
        for w0 in range(0, len(prompts), self.max_batch):
            wave = prompts[w0:w0 + self.max_batch]
            self.llm.prefill(wave)
            if on_progress is None:
                self.llm.decode(max_new_tokens - 1, use_graph=use_graph)
            else:
                done, chunk = 1, max(1, int(chunk))
                on_progress(w0, self.llm.tokens(done))
                while done < max_new_tokens:
                    n = min(chunk, max_new_tokens - done)
                    self.llm.decode(n, use_graph=use_graph)
                    done += n
                    if done < max_new_tokens:
                        on_progress(w0, self.llm.tokens(done))
            out[w0:w0 + len(wave)] = self.llm.tokens(max_new_tokens)
            if on_progress is not None:
                on_progress(w0, out[w0:w0 + len(wave)])
        return out

    # ---- continuous batching: one scheduler iteration (clearml_serving_b200/llm_service.ContinuousBatcher)
    def kv_info(self):
        return self.llm.kv_info()

    def step(self, page_updates=(), prefill=None, decode=None, use_graph=True):
        """One iteration of the host scheduler over the paged KV cache:
          page_updates  [(slot, first logical page, [pool pages])...]  page-table rows to extend first
          prefill       (prompts, slots): new sequences into free KV slots -> their first sampled tokens int32[n]
          decode        (slots, ctx_len, last_tok, n_steps): advance these sequences n_steps -> tokens int32[rows, n_steps]
        Sequences in slots not named keep their cache untouched."""
        for slot, first, pages in page_updates:
            self.llm.set_pages(slot, first, pages)
        first_tokens = new_tokens = None
        if prefill is not None:
            prompts, slots = prefill
            self.llm.prefill_slots(prompts, slots)
            first_tokens = self.llm.tokens(1)[:, 0].copy()
        if decode is not None:
            slots, ctx_len, last_tok, n_steps = decode
            self.llm.set_rows(slots, ctx_len, last_tok)
            self.llm.decode(int(n_steps), use_graph=use_graph)
            new_tokens = self.llm.tokens(int(n_steps))
        return first_tokens, new_tokens

    def close(self):
        self.llm.free()
