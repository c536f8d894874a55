"""Serving side of the LLM endpoint: the engine class behind the reference's OpenAI-compatible routes.

Mirrors `VllmPreprocessRequest` (clearml_serving/serving/preprocess_service.py:1097-1348): the REST layer calls
`getattr(engine, serve_type.replace("/", "_"))` for `POST {prefix}/openai/{serve_type}` (serving/main.py:217-231;
model_request_processor.py:1331), so the engine exposes `v1_completions`, `v1_chat_completions`, `v1_models`,
`tokenize`, `detokenize`, `version` with the reference's (data, state, collect_custom_statistics_fn) signature
and async flags.  What replaces vLLM underneath:

  * `ContinuousBatcher` -- per-endpoint scheduler with CONTINUOUS BATCHING over the PAGED KV cache (what vLLM's
                       scheduler + block manager do behind the reference's engine): a request joins the running
                       batch at the next scheduler iteration -- its prompt is prefilled into a free KV slot while the
                       other sequences keep their cache -- and leaves it the moment it finishes (max_tokens, EOS or a
                       stop token), freeing its slot and pages for the next arrival.  KV pages (64 tokens) are handed
                       out lazily from one pool as sequences grow; admission reserves the worst case so nothing is
                       ever preempted.  `"stream": true` requests get their tokens after every iteration as
                       server-sent events.
  * `WaveBatcher`   -- the static-wave scheduler of round 1 (`"b200.llm": {"scheduler": "waves"}`): collects up to
                       `max_batch` prompts, runs ONE prefill + CUDA-graph decode wave.
  * `TensorParallelLeader` / `follower_loop` -- with tensor_parallel_size 2 the serving process is rank 0; rank 1
                       is a worker process that replays every engine call (torchrun starts both,
                       `python -m torch.distributed.run --nproc-per-node 2 -m clearml_serving_b200.llm_service`).
                       Only the prompts travel over torch.distributed; activations go over NVLink peer memory.

Model configuration comes from the endpoint's `auxiliary_cfg` like the reference's vLLM engine args
(examples/vllm/preprocess.py): `{"b200.llm": {"architecture": {...LlamaConfig fields...} | "llama3-8b",
"load_format": "dummy" | "safetensors", "seed": 0, "max_batch": 32, "max_model_len": 1024,
"tensor_parallel_size": 1 | 2}}`.  Text prompts need a tokenizer object from the endpoint's user code
(`Preprocess.load()` may return `{"tokenizer": tok}`); token-id prompts -- which the OpenAI completions API
allows -- need none.
"""
import asyncio
import os
import threading
import time
import uuid
from concurrent.futures import Future

import numpy as np

from . import llm as L
from .preprocess_service import BasePreprocessRequest

__version__ = "0.1"


class _Request(object):
    __slots__ = ("prompt", "max_tokens", "future", "t_enqueue", "on_tokens", "sent", "stop_ids", "finish_reason",
                 "slot", "pages", "tokens", "ctx_len", "t_first")

    def __init__(self, prompt, max_tokens, on_tokens=None, stop_ids=()):
        self.prompt = prompt
        self.max_tokens = int(max_tokens)
        self.future = Future()
        self.t_enqueue = time.perf_counter()
        self.on_tokens = on_tokens   # streaming: called on the engine thread with (new token ids, finished)
        self.sent = 0
        self.stop_ids = frozenset(int(t) for t in stop_ids)   # EOS / stop_token_ids: generation ends AFTER such a token
        self.finish_reason = "length"
        self.slot, self.pages, self.tokens, self.ctx_len, self.t_first = -1, [], [], 0, None


class GenerationResult(np.ndarray):
    """the generated token ids (int32 array) + why generation ended ("length" | "stop")"""
    finish_reason = "length"


def _result(tokens, finish_reason):
    r = np.asarray(tokens, dtype=np.int32).view(GenerationResult)
    r.finish_reason = finish_reason
    return r


PAGE_TOKENS = 64


class ContinuousBatcher(object):
    """Iteration-level scheduler (continuous batching) over the engine's paged KV cache.  One thread owns the engine;
    `submit()` is thread-safe.  Per iteration: (1) admit queued requests into free KV slots while the page pool can
    cover their worst case, (2) prefill the newcomers (one ragged batch, other sequences untouched), (3) run up to
    `chunk` decode steps for every running sequence (never past the nearest max_tokens), (4) hand out the new tokens,
    retire finished sequences."""

    def __init__(self, engine, max_batch, max_ctx, name="llm", chunk=4, max_prefill_tokens=None):
        self.engine = engine
        self.max_batch, self.max_ctx = int(max_batch), int(max_ctx)
        self.chunk = max(1, int(chunk))
        self.n_pages, page_tokens, self.pages_per_seq = engine.kv_info()
        assert page_tokens == PAGE_TOKENS
        self.max_prefill_tokens = int(max_prefill_tokens or getattr(engine, "max_prefill_tokens", self.max_batch * self.max_ctx))
        self._free_slots = list(range(self.max_batch))
        self._free_pages = list(range(self.n_pages - 1, -1, -1))
        self._reserved = 0            # pages promised to running sequences (worst case), >= pages handed out
        self._active = []
        self._cv = threading.Condition()
        self._queue = []
        self._closed = False
        self.stats = dict(iterations=0, requests=0, prefill_batches=0, decode_steps=0, rows_sum=0, max_rows=0, joined_running=0,
                          queue_s=0.0, ttft_s=0.0, pages_peak=0, finished_stop=0)
        self._thread = threading.Thread(target=self._run, name="b2s-llm-" + name, daemon=True)
        self._thread.start()

    @staticmethod
    def _pages_for(n_tokens):
        return (int(n_tokens) + PAGE_TOKENS - 1) // PAGE_TOKENS

    def submit(self, prompt, max_tokens, on_tokens=None, stop_ids=()):
        if len(prompt) < 1 or len(prompt) + int(max_tokens) > self.max_ctx:
            raise ValueError("prompt ({} tokens) + max_tokens ({}) exceeds max_model_len {}".format(len(prompt), max_tokens, self.max_ctx))
        if len(prompt) > self.max_prefill_tokens:
            raise ValueError("prompt of {} tokens exceeds max_num_batched_tokens {}".format(len(prompt), self.max_prefill_tokens))
        if self._pages_for(len(prompt) + int(max_tokens)) > self.n_pages:
            raise ValueError("prompt + max_tokens need more KV pages than the pool holds ({})".format(self.n_pages))
        r = _Request(prompt, max_tokens, on_tokens, stop_ids)
        with self._cv:
            if self._closed:
                raise RuntimeError("llm endpoint is shutting down")
            self._queue.append(r)
            self._cv.notify()
        return r.future

    # ---- engine thread
    def _admit(self):
        new, budget = [], self.max_prefill_tokens
        with self._cv:
            while self._queue and self._free_slots:
                r = self._queue[0]
                need = self._pages_for(len(r.prompt) + r.max_tokens)
                if self._reserved + need > self.n_pages or len(r.prompt) > budget:
                    break                       # FIFO: the head waits for pages / the next prefill batch
                self._queue.pop(0)
                self._reserved += need
                budget -= len(r.prompt)
                r.slot = self._free_slots.pop(0)
                new.append(r)
        return new

    def _grow(self, r, upto_tokens, updates):
        """make sure the sequence owns pages for positions [0, upto_tokens)"""
        want = self._pages_for(upto_tokens)
        if want > len(r.pages):
            first = len(r.pages)
            fresh = [self._free_pages.pop() for _ in range(want - first)]   # cannot run dry: admission reserved them
            r.pages.extend(fresh)
            updates.append((r.slot, first, fresh))

    def _deliver(self, r, new_tokens):
        """append tokens, honour stop tokens / max_tokens; returns True when the sequence is finished"""
        done = False
        for t in new_tokens:
            r.tokens.append(int(t))
            if int(t) in r.stop_ids:
                r.finish_reason, done = "stop", True
                break
            if len(r.tokens) >= r.max_tokens:
                done = True
                break
        if r.on_tokens is not None and len(r.tokens) > r.sent:
            r.on_tokens(r.tokens[r.sent:], done)
            r.sent = len(r.tokens)
        return done

    def _retire(self, r):
        with self._cv:
            self._free_slots.append(r.slot)
            self._free_pages.extend(reversed(r.pages))
            self._reserved -= self._pages_for(len(r.prompt) + r.max_tokens)
        if r.finish_reason == "stop":
            self.stats["finished_stop"] += 1
        r.pages = []
        if not r.future.done():
            r.future.set_result(_result(r.tokens, r.finish_reason))

    def _run(self):
        while True:
            with self._cv:
                while not self._queue and not self._active and not self._closed:
                    self._cv.wait()
                if self._closed and not self._queue and not self._active:
                    return
            new = self._admit()
            if not new and not self._active:
                with self._cv:              # queue head does not fit yet and nothing runs: cannot happen unless closing
                    if self._closed:
                        for r in self._queue:
                            r.future.set_exception(RuntimeError("llm endpoint shut down"))
                        self._queue = []
                        return
                    self._cv.wait(0.01)
                continue
            st = self.stats
            batch = new + self._active
            try:
                updates, prefill, decode = [], None, None
                t0 = time.perf_counter()
                if new:
                    for r in new:
                        self._grow(r, len(r.prompt) + 1, updates)
                    prefill = ([r.prompt for r in new], [r.slot for r in new])
                    if self._active:
                        st["joined_running"] += len(new)
                n_steps = 0
                if self._active:
                    n_steps = min([self.chunk] + [r.max_tokens - len(r.tokens) for r in self._active])
                    for r in self._active:
                        self._grow(r, r.ctx_len + n_steps + 1, updates)
                    decode = ([r.slot for r in self._active], [r.ctx_len for r in self._active], [r.tokens[-1] for r in self._active], n_steps)
                first, toks = self.engine.step(updates, prefill, decode)
                now = time.perf_counter()
                st["iterations"] += 1
                st["pages_peak"] = max(st["pages_peak"], self.n_pages - len(self._free_pages))
                still = []
                if decode is not None:
                    st["decode_steps"] += n_steps
                    st["rows_sum"] += n_steps * len(self._active)
                    st["max_rows"] = max(st["max_rows"], len(self._active))
                    for i, r in enumerate(self._active):
                        r.ctx_len += n_steps
                        if self._deliver(r, toks[i]):
                            self._retire(r)
                        else:
                            still.append(r)
                if prefill is not None:
                    st["prefill_batches"] += 1
                    st["requests"] += len(new)
                    for i, r in enumerate(new):
                        r.ctx_len, r.t_first = len(r.prompt), now
                        st["queue_s"] += t0 - r.t_enqueue
                        st["ttft_s"] += now - r.t_enqueue
                        if self._deliver(r, [first[i]]):
                            self._retire(r)
                        else:
                            still.append(r)
                self._active = still
            except Exception as ex:  # noqa -- an engine error fails the sequences of this iteration (-> 422 / restart upstream)
                for r in batch:
                    if not r.future.done():
                        r.future.set_exception(ex)
                with self._cv:
                    for r in batch:
                        if r.slot >= 0 and r.slot not in self._free_slots:
                            self._free_slots.append(r.slot)
                            self._free_pages.extend(reversed(r.pages))
                            self._reserved -= self._pages_for(len(r.prompt) + r.max_tokens)
                            r.pages = []
                self._active = []

    def close(self):
        with self._cv:
            self._closed = True
            self._cv.notify_all()
        self._thread.join(timeout=60)


class WaveBatcher(object):
    """Timeout / max-batch scheduler for generation requests (the Triton dynamic-batcher keys of
    `triton_helper.create_config_pbtxt`, triton_helper.py:291-409, applied to prompt waves)."""

    def __init__(self, engine, max_batch, max_queue_delay_us=2000, name="llm", stream_chunk=8):
        self.engine = engine
        self.stream_chunk = int(stream_chunk)   # decode steps between two deliveries to streaming clients
        self.max_batch = int(max_batch)
        self.delay_s = max(0.0, float(max_queue_delay_us) * 1e-6)
        self._cv = threading.Condition()
        self._queue = []
        self._closed = False
        self.stats = dict(waves=0, requests=0, wave_sizes={}, queue_s=0.0, engine_s=0.0)
        self._thread = threading.Thread(target=self._run, name="b2s-llm-" + name, daemon=True)
        self._thread.start()

    def submit(self, prompt, max_tokens, on_tokens=None, stop_ids=()):
        r = _Request(prompt, max_tokens, on_tokens, stop_ids)
        with self._cv:
            if self._closed:
                raise RuntimeError("llm endpoint is shutting down")
            self._queue.append(r)
            self._cv.notify()
        return r.future

    def _take_wave(self):
        with self._cv:
            while not self._queue and not self._closed:
                self._cv.wait()
            if self._closed and not self._queue:
                return None
            deadline = self._queue[0].t_enqueue + self.delay_s
            while len(self._queue) < self.max_batch and not self._closed:
                left = deadline - time.perf_counter()
                if left <= 0:
                    break
                self._cv.wait(left)
            # a wave generates max(max_tokens) tokens for EVERY member: admit a request only while each member's
            # prompt + that maximum still fits the context (otherwise one client's long budget fails the others)
            max_ctx = getattr(getattr(self.engine, "engine", self.engine), "max_ctx", None)
            wave, n_new = [], 0
            for r in self._queue[:self.max_batch]:
                cand = max(n_new, r.max_tokens)
                if max_ctx is not None and wave and any(len(q.prompt) + cand > max_ctx for q in wave + [r]):
                    break
                wave.append(r)
                n_new = cand
            self._queue = self._queue[len(wave):]
            return wave

    def _run(self):
        while True:
            wave = self._take_wave()
            if wave is None:
                return
            t0 = time.perf_counter()
            try:
                n_new = max(r.max_tokens for r in wave)
                if any(r.on_tokens is not None for r in wave):
                    def progress(w0, toks, wave=wave):   # engine thread: hand every streaming caller its new tokens
                        for i, r in enumerate(wave[w0:w0 + len(toks)]):
                            upto = min(toks.shape[1], r.max_tokens)
                            if r.on_tokens is not None and upto > r.sent:
                                r.on_tokens([int(t) for t in toks[i, r.sent:upto]], upto == r.max_tokens)
                                r.sent = upto
                    out = self.engine.generate([r.prompt for r in wave], n_new, on_progress=progress, chunk=self.stream_chunk)
                else:
                    out = self.engine.generate([r.prompt for r in wave], n_new)
                for i, r in enumerate(wave):
                    toks = [int(t) for t in out[i, :r.max_tokens]]
                    reason = "length"
                    for k, t in enumerate(toks):
                        if t in r.stop_ids:
                            toks, reason = toks[:k + 1], "stop"
                            break
                    r.future.set_result(_result(toks, reason))
            except Exception as ex:  # noqa -- every caller of the wave sees the engine error (mapped to 422 / restart)
                for r in wave:
                    if not r.future.done():
                        r.future.set_exception(ex)
            t1 = time.perf_counter()
            st = self.stats
            st["waves"] += 1
            st["requests"] += len(wave)
            st["wave_sizes"][len(wave)] = st["wave_sizes"].get(len(wave), 0) + 1
            st["queue_s"] += sum(t0 - r.t_enqueue for r in wave)
            st["engine_s"] += t1 - t0

    def close(self):
        with self._cv:
            self._closed = True
            self._cv.notify_all()
        self._thread.join(timeout=30)


# ------------------------------------------------------------------------------------------------
# tensor-parallel pair: rank 0 leads, rank 1 replays
# ------------------------------------------------------------------------------------------------
class TensorParallelLeader(object):
    """Wraps an LlmEngine on rank 0 of a 2-rank group: every call is announced to the follower first, so both
    ranks issue the same kernel sequence (the data path itself never touches torch.distributed)."""

    def __init__(self, engine, group=None):
        self.engine, self.group = engine, group

    def _announce(self, msg):
        import torch.distributed as dist
        dist.broadcast_object_list([msg], src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)

    def generate(self, prompts, max_new_tokens, on_progress=None, chunk=8):
        # the follower chunks its decode loop the same way, so both ranks synchronise at the same steps
        check = getattr(self.engine, "_check", None)
        if callable(check):
            check(prompts, int(max_new_tokens))   # refuse BEFORE announcing: the follower never sees a doomed call
        self._announce(("generate", [np.asarray(p, np.int32) for p in prompts], int(max_new_tokens),
                        int(chunk) if on_progress is not None else 0))
        return self.engine.generate(prompts, max_new_tokens, on_progress=on_progress, chunk=chunk)

    def kv_info(self):
        return self.engine.kv_info()

    @property
    def max_prefill_tokens(self):
        return self.engine.max_prefill_tokens

    def step(self, page_updates=(), prefill=None, decode=None):
        # validated here so that an error cannot leave the follower waiting inside a collective-free replay
        self._announce(("step", list(page_updates), prefill, decode))
        return self.engine.step(page_updates, prefill, decode)

    def close(self):
        self._announce(("close",))
        self.engine.close()


def follower_loop(engine, group=None):
    import torch.distributed as dist
    src = dist.get_global_rank(group, 0) if group is not None else 0
    while True:
        box = [None]
        dist.broadcast_object_list(box, src=src, group=group)
        msg = box[0]
        if msg[0] == "generate":
            chunk = msg[3] if len(msg) > 3 else 0
            try:   # both ranks stay in lock-step even when the call is refused (the leader raised the same error)
                engine.generate(msg[1], msg[2], on_progress=(lambda w0, toks: None) if chunk else None, chunk=chunk or 8)
            except ValueError:
                pass
        elif msg[0] == "step":
            try:
                engine.step(msg[1], msg[2], msg[3])
            except ValueError:
                pass
        elif msg[0] == "close":
            engine.close()
            return


# ------------------------------------------------------------------------------------------------
# engine construction from auxiliary_cfg
# ------------------------------------------------------------------------------------------------
def _spec_from_cfg(cfg, model_path=None):
    arch = cfg.get("architecture", None)
    if arch in (None, "auto") and model_path:
        import json
        cj = os.path.join(model_path, "config.json") if os.path.isdir(model_path) else None
        if cj and os.path.exists(cj):
            with open(cj) as f:
                arch = json.load(f)
    if arch in ("llama3-8b", "meta-llama/Meta-Llama-3-8B"):
        return L.LlamaSpec.llama3_8b()
    if not isinstance(arch, dict):
        raise ValueError("b200 llm engine: `b200.llm.architecture` must be 'llama3-8b' or a dict of LlamaConfig fields")
    hd = arch.get("head_dim") or arch["hidden_size"] // arch["num_attention_heads"]
    theta = arch.get("rope_theta") or (arch.get("rope_parameters") or {}).get("rope_theta", 10000.0)
    return L.LlamaSpec(vocab_size=arch["vocab_size"], hidden_size=arch["hidden_size"], intermediate_size=arch["intermediate_size"],
                       num_hidden_layers=arch["num_hidden_layers"], num_attention_heads=arch["num_attention_heads"],
                       num_key_value_heads=arch.get("num_key_value_heads", arch["num_attention_heads"]), head_dim=hd,
                       rope_theta=float(theta), rms_norm_eps=float(arch.get("rms_norm_eps", 1e-5)))


def _load_safetensors(path):
    from safetensors import safe_open   # ships with transformers
    files = [path] if os.path.isfile(path) else sorted(
        os.path.join(path, f) for f in os.listdir(path) if f.endswith(".safetensors"))
    if not files:
        raise ValueError("b200 llm engine: no .safetensors file under {}".format(path))
    sd = {}
    for fn in files:
        with safe_open(fn, framework="pt") as f:
            for k in f.keys():
                sd[k] = f.get_tensor(k)
    return sd


def build_engine(cfg, model_path=None, device=0, tp_rank=0, tp_group=None):
    """cfg: the `b200.llm` dict of the endpoint's auxiliary_cfg -> LlmEngine (weights loaded / initialised)"""
    spec = _spec_from_cfg(cfg, model_path)
    tp = int(cfg.get("tensor_parallel_size", 1))
    max_batch = int(cfg.get("max_batch", cfg.get("max_num_seqs", 32)))
    max_len = int(cfg.get("max_model_len", 1024))
    # paged KV pool: `kv_pages` (pages of 64 tokens) or vLLM's `num_gpu_blocks_override`; 0 = every slot can reach max_model_len
    kv_pages = int(cfg.get("kv_pages", cfg.get("num_gpu_blocks_override", 0)) or 0)
    eng = L.LlmEngine(spec, device=device, max_batch=max_batch, max_ctx=max_len,
                      max_tokens=int(cfg.get("max_num_batched_tokens", max_batch * max_len)), tp_size=tp, tp_rank=tp_rank,
                      tp_group=tp_group, kv_pages=kv_pages)
    fmt = cfg.get("load_format", "safetensors" if model_path else "dummy")
    if fmt == "dummy":   # vLLM's name for random weights (BASELINE.json configs[4])
        eng.init_random(seed=int(cfg.get("seed", 0)), std=float(cfg.get("init_std", 0.02)))
    elif fmt == "safetensors":
        if not model_path:
            raise ValueError("b200 llm engine: load_format=safetensors needs a model path")
        eng.load_state_dict(_load_safetensors(model_path))
    else:
        raise ValueError("b200 llm engine: unknown load_format '{}'".format(fmt))
    return eng


@BasePreprocessRequest.register_engine("b200_llm", modules=["numpy"])
class B200LlmPreprocessRequest(BasePreprocessRequest):
    is_preprocess_async = True
    is_process_async = True
    is_postprocess_async = True

    def __init__(self, model_endpoint, task=None):
        super(B200LlmPreprocessRequest, self).__init__(model_endpoint=model_endpoint, task=task)
        aux = getattr(model_endpoint, "auxiliary_cfg", None)
        cfg = dict((aux or {}).get("b200.llm", {})) if isinstance(aux, dict) else {}
        user = self._model if isinstance(self._model, dict) else {}
        self._tokenizer = user.get("tokenizer")
        path = None
        if getattr(model_endpoint, "model_id", None):
            path = self._get_local_model_file()
        tp = int(cfg.get("tensor_parallel_size", 1))
        device = int(cfg.get("device", os.environ.get("LOCAL_RANK", 0)))
        engine = build_engine(cfg, model_path=path, device=device, tp_rank=0)
        self._engine = TensorParallelLeader(engine) if tp == 2 else engine
        self._spec = engine.spec
        self._max_ctx = engine.max_ctx
        delay = 2000
        if isinstance(aux, dict):
            delay = aux.get("dynamic_batching", {}).get("max_queue_delay_microseconds", delay) if isinstance(
                aux.get("dynamic_batching"), dict) else aux.get("dynamic_batching.max_queue_delay_microseconds", delay)
        if str(cfg.get("scheduler", "continuous")) == "waves":
            self._batcher = WaveBatcher(self._engine, engine.max_batch, delay, name=str(model_endpoint.serving_url))
        else:
            self._batcher = ContinuousBatcher(self._engine, engine.max_batch, engine.max_ctx, name=str(model_endpoint.serving_url),
                                              chunk=int(cfg.get("decode_chunk", 4)))
        self._model_name = str(model_endpoint.serving_url)
        self._eos_ids = self._eos_from(cfg, self._tokenizer)

    # ---- reference plugin surface (async pass-throughs unless the user code overrides them)
    async def preprocess(self, request, state, collect_custom_statistics_fn=None):
        if self._preprocess is not None and hasattr(self._preprocess, "preprocess"):
            r = self._preprocess.preprocess(request, state, collect_custom_statistics_fn)
            return await r if asyncio.iscoroutine(r) else r
        return request

    async def postprocess(self, data, state, collect_custom_statistics_fn=None):
        if self._preprocess is not None and hasattr(self._preprocess, "postprocess"):
            r = self._preprocess.postprocess(data, state, collect_custom_statistics_fn)
            return await r if asyncio.iscoroutine(r) else r
        return data

    @staticmethod
    def _body(data):
        body = data.get("request", data) if isinstance(data, dict) else data
        if hasattr(body, "model_dump"):
            body = body.model_dump()
        if not isinstance(body, dict):
            raise ValueError("llm request body must be a JSON object")
        return body

    def _encode(self, prompt):
        if isinstance(prompt, str):
            if self._tokenizer is None:
                raise ValueError("text prompts need a tokenizer (Preprocess.load() -> {'tokenizer': ...}); send token ids instead")
            ids = self._tokenizer.encode(prompt)
        else:
            ids = prompt
        ids = np.asarray(ids, dtype=np.int64).reshape(-1)
        if ids.size < 1 or ids.min() < 0 or ids.max() >= self._spec.vocab_size:
            raise ValueError("prompt token ids must be in [0, {})".format(self._spec.vocab_size))
        return ids.astype(np.int32)

    @staticmethod
    def _eos_from(cfg, tokenizer):
        ids = cfg.get("eos_token_id", getattr(tokenizer, "eos_token_id", None) if tokenizer is not None else None)
        if ids is None:
            return ()
        return tuple(int(t) for t in (ids if isinstance(ids, (list, tuple)) else [ids]))

    def _sampling(self, body):
        """OpenAI / vLLM sampling fields -> the stop-token set of a greedy generation.  What the engine cannot honour is
        REFUSED (ValueError -> 422) instead of silently answered with different output."""
        temp = body.get("temperature")
        if temp not in (None, 0, 0.0):
            raise ValueError("b200 llm engine decodes greedily: temperature must be 0 (got {})".format(temp))
        if body.get("n") not in (None, 1) or body.get("best_of") not in (None, 1):
            raise ValueError("b200 llm engine: n / best_of > 1 are not supported")
        for k in ("top_p", "top_k", "presence_penalty", "frequency_penalty", "repetition_penalty", "logit_bias", "logprobs"):
            v = body.get(k)
            if v not in (None, 0, 0.0, 1, 1.0, -1, {}, False) or (k in ("top_p", "repetition_penalty") and v not in (None, 1, 1.0)) and v is not None:
                raise ValueError("b200 llm engine decodes greedily: '{}' is not supported".format(k))
        if body.get("stop"):
            raise ValueError("b200 llm engine: 'stop' strings are not supported (use stop_token_ids)")
        stop = set(int(t) for t in (body.get("stop_token_ids") or []))
        if not body.get("ignore_eos"):
            stop.update(self._eos_ids)
        return stop

    def _check_len(self, prompts, max_tokens):
        for p in prompts:
            if len(p) + max_tokens > self._max_ctx:
                raise ValueError("prompt ({} tokens) + max_tokens ({}) exceeds max_model_len {}".format(len(p), max_tokens, self._max_ctx))

    def _stream(self, prompts, max_tokens, make_chunk, stop_ids=()):
        """Server-sent events for `"stream": true` (the reference hands vLLM's generator to a StreamingResponse,
        preprocess_service.py:1219-1234 / :1262-1277): one `data: {chunk}` event per delivery of the wave's decode loop and
        per prompt, `finish_reason` on a prompt's last one, then `data: [DONE]`.  `make_chunk(index, token_ids, text,
        finish_reason)` builds the OpenAI chunk object; text is the newly completed suffix of the decoded output."""
        import json
        from starlette.responses import StreamingResponse
        self._check_len(prompts, max_tokens)
        loop = asyncio.get_running_loop()
        queue = asyncio.Queue()
        futs = []
        for i, p in enumerate(prompts):
            def on_tokens(toks, finished, i=i):
                loop.call_soon_threadsafe(queue.put_nowait, (i, toks, finished, None))
            f = self._batcher.submit(p, max_tokens, on_tokens, stop_ids)
            f.add_done_callback(lambda f, i=i: f.exception() is not None and
                                loop.call_soon_threadsafe(queue.put_nowait, (i, [], True, f.exception())))
            futs.append(f)
        reasons = {}
        for i, f in enumerate(futs):
            f.add_done_callback(lambda f, i=i: f.exception() is None and reasons.__setitem__(i, getattr(f.result(), "finish_reason", "length")))

        async def events():
            open_prompts, seen, texts = len(prompts), [[] for _ in prompts], ["" for _ in prompts]
            while open_prompts:
                i, toks, finished, err = await queue.get()
                if err is not None:
                    yield "data: " + json.dumps(dict(error=dict(message=str(err), type=type(err).__name__))) + "\n\n"
                    break
                seen[i].extend(toks)
                text = ""
                if self._tokenizer is not None:
                    full = self._tokenizer.decode(seen[i])
                    if finished or not full.endswith("\ufffd"):   # hold back an incomplete multi-token character
                        text, texts[i] = full[len(texts[i]):], full
                if finished and i not in reasons:       # the future resolves right after the last delivery
                    try:
                        reasons[i] = getattr(await asyncio.wait_for(asyncio.wrap_future(futs[i]), 5.0), "finish_reason", "length")
                    except Exception:  # noqa
                        reasons[i] = "length"
                yield "data: " + json.dumps(make_chunk(i, toks, text, reasons.get(i, "length") if finished else None)) + "\n\n"
                open_prompts -= 1 if finished else 0
            yield "data: [DONE]\n\n"
        return StreamingResponse(content=events(), media_type="text/event-stream")

    async def _generate(self, prompts, max_tokens, stop_ids=()):
        self._check_len(prompts, max_tokens)
        futs = [asyncio.wrap_future(self._batcher.submit(p, max_tokens, None, stop_ids)) for p in prompts]
        return await asyncio.gather(*futs)

    async def v1_completions(self, data, state, collect_custom_statistics_fn=None):
        body = self._body(data)
        prompt = body.get("prompt")
        if prompt is None:
            raise ValueError("completions: 'prompt' is required")
        many = isinstance(prompt, (list, tuple)) and len(prompt) > 0 and isinstance(prompt[0], (list, tuple, str))
        prompts = [self._encode(p) for p in (prompt if many else [prompt])]
        max_tokens = 16 if body.get("max_tokens") is None else int(body["max_tokens"])   # OpenAI default
        if max_tokens < 1:
            raise ValueError("completions: max_tokens must be >= 1")
        stop_ids = self._sampling(body)
        if body.get("stream"):
            cid, created, model = "cmpl-" + uuid.uuid4().hex, int(time.time()), body.get("model") or self._model_name
            return self._stream(prompts, max_tokens, lambda i, toks, text, fin: dict(
                id=cid, object="text_completion", created=created, model=model,
                choices=[dict(index=i, text=text, token_ids=toks, logprobs=None, finish_reason=fin)]), stop_ids)
        outs = await self._generate(prompts, max_tokens, stop_ids)
        choices = []
        for i, res in enumerate(outs):
            toks = [int(t) for t in res]
            text = self._tokenizer.decode(toks) if self._tokenizer is not None else ""
            choices.append(dict(index=i, text=text, token_ids=toks, logprobs=None, finish_reason=getattr(res, "finish_reason", "length")))
        n_prompt = int(sum(len(p) for p in prompts))
        n_out = int(sum(len(c["token_ids"]) for c in choices))
        if collect_custom_statistics_fn:
            collect_custom_statistics_fn(dict(prompt_tokens=n_prompt, completion_tokens=n_out))
        return dict(id="cmpl-" + uuid.uuid4().hex, object="text_completion", created=int(time.time()), model=body.get("model") or self._model_name,
                    choices=choices, usage=dict(prompt_tokens=n_prompt, completion_tokens=n_out, total_tokens=n_prompt + n_out))

    async def v1_chat_completions(self, data, state, collect_custom_statistics_fn=None):
        body = self._body(data)
        if self._tokenizer is None or not hasattr(self._tokenizer, "apply_chat_template"):
            raise ValueError("chat completions need a tokenizer with a chat template (Preprocess.load() -> {'tokenizer': ...})")
        ids = self._tokenizer.apply_chat_template(body.get("messages") or [], add_generation_prompt=True, tokenize=True)
        mt = body.get("max_tokens") if body.get("max_tokens") is not None else body.get("max_completion_tokens")
        max_tokens = 16 if mt is None else int(mt)
        if max_tokens < 1:
            raise ValueError("chat completions: max_tokens must be >= 1")
        stop_ids = self._sampling(body)
        if body.get("stream"):
            cid, created, model = "chatcmpl-" + uuid.uuid4().hex, int(time.time()), body.get("model") or self._model_name
            first = [True]

            def chunk(i, toks, text, fin):
                delta = dict(content=text)
                if first[0]:
                    delta["role"], first[0] = "assistant", False
                return dict(id=cid, object="chat.completion.chunk", created=created, model=model,
                            choices=[dict(index=0, delta=delta, token_ids=toks, finish_reason=fin)])
            return self._stream([self._encode(ids)], max_tokens, chunk, stop_ids)
        res = (await self._generate([self._encode(ids)], max_tokens, stop_ids))[0]
        toks = [int(t) for t in res]
        msg = dict(role="assistant", content=self._tokenizer.decode(toks))
        return dict(id="chatcmpl-" + uuid.uuid4().hex, object="chat.completion", created=int(time.time()),
                    model=body.get("model") or self._model_name,
                    choices=[dict(index=0, message=msg, token_ids=toks, finish_reason=getattr(res, "finish_reason", "length"))],
                    usage=dict(prompt_tokens=len(ids), completion_tokens=len(toks), total_tokens=len(ids) + len(toks)))

    async def v1_models(self, data, state, collect_custom_statistics_fn=None):
        return dict(object="list", data=[dict(id=self._model_name, object="model", owned_by="b200serve", max_model_len=self._max_ctx)])

    async def version(self, data, state, collect_custom_statistics_fn=None):
        return dict(version="b200serve-llm " + __version__)

    async def tokenize(self, data, state, collect_custom_statistics_fn=None):
        ids = [int(t) for t in self._encode(self._body(data).get("prompt", ""))]
        return dict(tokens=ids, count=len(ids), max_model_len=self._max_ctx)

    async def detokenize(self, data, state, collect_custom_statistics_fn=None):
        if self._tokenizer is None:
            raise ValueError("detokenize needs a tokenizer (Preprocess.load() -> {'tokenizer': ...})")
        return dict(prompt=self._tokenizer.decode(list(self._body(data).get("tokens") or [])))

    async def process(self, data, state, collect_custom_statistics_fn=None):
        # plain POST /serve/<endpoint>: same contract as /openai/v1/completions
        return await self.v1_completions(data, state, collect_custom_statistics_fn)

    def engine_stats(self):
        return dict(self._batcher.stats)

    def unload(self):
        b, self._batcher = getattr(self, "_batcher", None), None
        if b is not None:
            b.close()
        e, self._engine = getattr(self, "_engine", None), None
        if e is not None:
            e.close()

    def __del__(self):
        try:
            self.unload()
        except Exception:  # noqa
            pass


def main():
    """torchrun entry for a tensor-parallel pair: rank 0 serves REST (uvicorn), rank 1 replays engine calls.
    Endpoint configuration as for the single-process server (CLEARML_SERVING_* / b200serve config file)."""
    import json
    import torch.distributed as dist
    dist.init_process_group("gloo")
    rank = dist.get_rank()
    if rank == 0:
        import uvicorn
        from .main import create_app
        uvicorn.run(create_app(), host=os.environ.get("B2S_HOST", "0.0.0.0"), port=int(os.environ.get("B2S_PORT", "8080")))
    else:
        cfg = json.loads(os.environ.get("B2S_LLM_CFG", "{}"))
        cfg["tensor_parallel_size"] = 2
        follower_loop(build_engine(cfg, model_path=os.environ.get("B2S_LLM_MODEL_PATH") or None,
                                   device=int(os.environ.get("LOCAL_RANK", 1)), tp_rank=1))


if __name__ == "__main__":
    main()
