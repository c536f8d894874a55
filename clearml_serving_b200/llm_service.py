"""Serving side of the LLM endpoint: the engine class behind the reference's OpenAI-compatible routes.

Mirrors `VllmPreprocessRequest` (clearml_serving/serving/preprocess_service.py:1097-1348): the REST layer calls
`getattr(engine, serve_type.replace("/", "_"))` for `POST {prefix}/openai/{serve_type}` (serving/main.py:217-231;
model_request_processor.py:1331), so the engine exposes `v1_completions`, `v1_chat_completions`, `v1_models`,
`tokenize`, `detokenize`, `version` with the reference's (data, state, collect_custom_statistics_fn) signature
and async flags.  What replaces vLLM underneath:

  * `WaveBatcher`   -- per-endpoint request queue: collects up to `max_batch` prompts (or until
                       `max_queue_delay_microseconds` after the first one), runs ONE prefill + CUDA-graph decode
                       wave on the engine thread, completes the callers' futures; `"stream": true` requests get
                       their tokens every few decode steps as server-sent events.  (Static waves; continuous
                       batching is the next item of SURVEY.md section 8f.)
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
    __slots__ = ("prompt", "max_tokens", "future", "t_enqueue", "on_tokens", "sent")

    def __init__(self, prompt, max_tokens, on_tokens=None):
        self.prompt = prompt
        self.max_tokens = int(max_tokens)
        self.future = Future()
        self.t_enqueue = time.perf_counter()
        self.on_tokens = on_tokens   # streaming: called on the engine thread with (new token ids, finished)
        self.sent = 0


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

    def submit(self, prompt, max_tokens, on_tokens=None):
        r = _Request(prompt, max_tokens, on_tokens)
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
            wave, self._queue = self._queue[:self.max_batch], self._queue[self.max_batch:]
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
                    r.future.set_result(np.array(out[i, :r.max_tokens]))
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
        self._announce(("generate", [np.asarray(p, np.int32) for p in prompts], int(max_new_tokens),
                        int(chunk) if on_progress is not None else 0))
        return self.engine.generate(prompts, max_new_tokens, on_progress=on_progress, chunk=chunk)

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
            engine.generate(msg[1], msg[2], on_progress=(lambda w0, toks: None) if chunk else None, chunk=chunk or 8)
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
    eng = L.LlmEngine(spec, device=device, max_batch=max_batch, max_ctx=max_len,
                      max_tokens=int(cfg.get("max_num_batched_tokens", max_batch * max_len)), tp_size=tp, tp_rank=tp_rank,
                      tp_group=tp_group)
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
        self._batcher = WaveBatcher(self._engine, engine.max_batch, delay, name=str(model_endpoint.serving_url))
        self._model_name = str(model_endpoint.serving_url)

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

    def _check_len(self, prompts, max_tokens):
        for p in prompts:
            if len(p) + max_tokens > self._max_ctx:
                raise ValueError("prompt ({} tokens) + max_tokens ({}) exceeds max_model_len {}".format(len(p), max_tokens, self._max_ctx))

    def _stream(self, prompts, max_tokens, make_chunk):
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
            f = self._batcher.submit(p, max_tokens, on_tokens)
            f.add_done_callback(lambda f, i=i: f.exception() is not None and
                                loop.call_soon_threadsafe(queue.put_nowait, (i, [], True, f.exception())))
            futs.append(f)

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
                yield "data: " + json.dumps(make_chunk(i, toks, text, "length" if finished else None)) + "\n\n"
                open_prompts -= 1 if finished else 0
            yield "data: [DONE]\n\n"
        return StreamingResponse(content=events(), media_type="text/event-stream")

    async def _generate(self, prompts, max_tokens):
        self._check_len(prompts, max_tokens)
        futs = [asyncio.wrap_future(self._batcher.submit(p, max_tokens)) for p in prompts]
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
        if body.get("stream"):
            cid, created, model = "cmpl-" + uuid.uuid4().hex, int(time.time()), body.get("model") or self._model_name
            return self._stream(prompts, max_tokens, lambda i, toks, text, fin: dict(
                id=cid, object="text_completion", created=created, model=model,
                choices=[dict(index=i, text=text, token_ids=toks, logprobs=None, finish_reason=fin)]))
        outs = await self._generate(prompts, max_tokens)
        choices = []
        for i, toks in enumerate(outs):
            toks = [int(t) for t in toks]
            text = self._tokenizer.decode(toks) if self._tokenizer is not None else ""
            choices.append(dict(index=i, text=text, token_ids=toks, logprobs=None, finish_reason="length"))
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
        max_tokens = int(body.get("max_tokens") or body.get("max_completion_tokens") or 16)
        if body.get("stream"):
            cid, created, model = "chatcmpl-" + uuid.uuid4().hex, int(time.time()), body.get("model") or self._model_name
            first = [True]

            def chunk(i, toks, text, fin):
                delta = dict(content=text)
                if first[0]:
                    delta["role"], first[0] = "assistant", False
                return dict(id=cid, object="chat.completion.chunk", created=created, model=model,
                            choices=[dict(index=0, delta=delta, token_ids=toks, finish_reason=fin)])
            return self._stream([self._encode(ids)], max_tokens, chunk)
        toks = [int(t) for t in (await self._generate([self._encode(ids)], max_tokens))[0]]
        msg = dict(role="assistant", content=self._tokenizer.decode(toks))
        return dict(id="chatcmpl-" + uuid.uuid4().hex, object="chat.completion", created=int(time.time()),
                    model=body.get("model") or self._model_name, choices=[dict(index=0, message=msg, token_ids=toks, finish_reason="length")],
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
