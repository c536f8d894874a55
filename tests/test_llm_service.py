"""CPU tests of the LLM endpoint's serving side: the wave scheduler and the OpenAI-shaped engine surface
(reference: VllmPreprocessRequest, clearml_serving/serving/preprocess_service.py:1097-1348; route
serving/main.py:217-231).  The engine underneath is a host double -- no GPU, no product fallback."""
import asyncio
import threading
import time

import numpy as np
import pytest

from clearml_serving_b200 import llm_service as S
from clearml_serving_b200.endpoints import ModelEndpoint
from clearml_serving_b200.model_request_processor import ModelRequestProcessor
from clearml_serving_b200.preprocess_service import BasePreprocessRequest


class FakeLlm(object):
    """generate() -> token j of sequence i = (sum(prompt_i) + j) % 1000; records wave sizes"""

    def __init__(self, max_batch=4, max_ctx=64, latency_s=0.0, vocab=1000):
        self.max_batch, self.max_ctx, self.latency_s = max_batch, max_ctx, latency_s
        self.waves = []
        self.closed = False

        class _Spec(object):
            vocab_size = vocab
        self.spec = _Spec()

    def generate(self, prompts, n, on_progress=None, chunk=8):
        assert len(prompts) <= self.max_batch
        self.waves.append((len(prompts), n))
        time.sleep(self.latency_s)
        out = np.array([[(int(np.sum(p)) + j) % 1000 for j in range(n)] for p in prompts], dtype=np.int32)
        if on_progress is not None:   # like LlmEngine.generate: after the prefill, then every `chunk` steps, then the end
            self.progress_calls = 0
            for done in [1] + list(range(1 + chunk, n, chunk)) + [n]:
                on_progress(0, out[:, :done])
                self.progress_calls += 1
        return out

    def close(self):
        self.closed = True


def test_wave_batcher_groups_and_truncates():
    eng = FakeLlm(max_batch=4, latency_s=0.02)
    b = S.WaveBatcher(eng, max_batch=4, max_queue_delay_us=30000)
    try:
        futs = [b.submit(np.array([i, 1]), 2 + i % 3) for i in range(10)]
        outs = [f.result(timeout=10) for f in futs]
        for i, o in enumerate(outs):
            assert o.tolist() == [(i + 1 + j) % 1000 for j in range(2 + i % 3)]
        assert sum(n for n, _ in eng.waves) == 10 and max(n for n, _ in eng.waves) <= 4
        assert len(eng.waves) <= 4                         # 10 requests in waves of <= 4, not 10 waves of 1
        assert b.stats["requests"] == 10 and b.stats["waves"] == len(eng.waves)
    finally:
        b.close()


def test_wave_batcher_timeout_dispatches_partial_wave_and_propagates_errors():
    eng = FakeLlm(max_batch=8)
    b = S.WaveBatcher(eng, max_batch=8, max_queue_delay_us=20000)
    try:
        t0 = time.perf_counter()
        assert b.submit(np.array([5]), 3).result(timeout=5).tolist() == [5, 6, 7]
        assert time.perf_counter() - t0 < 1.0 and eng.waves == [(1, 3)]

        def boom(prompts, n, **kw):
            raise ValueError("CUDA out of memory. injected")
        eng.generate = boom
        with pytest.raises(ValueError, match="CUDA out of memory. "):     # the text the reference restarts on
            b.submit(np.array([1]), 1).result(timeout=5)
    finally:
        b.close()
    with pytest.raises(RuntimeError):
        b.submit(np.array([1]), 1)


def _engine(monkeypatch, tokenizer=None, max_batch=4, scheduler="waves", fake=None, extra=None):
    fake = fake or FakeLlm(max_batch=max_batch)
    monkeypatch.setattr(S, "build_engine", lambda cfg, model_path=None, device=0, tp_rank=0, tp_group=None: fake)
    cfg = {"architecture": "llama3-8b", "load_format": "dummy", "max_batch": max_batch, "scheduler": scheduler}
    cfg.update(extra or {})
    ep = ModelEndpoint(engine_type="b200_llm", serving_url="llama", auxiliary_cfg={
        "b200.llm": cfg, "dynamic_batching": {"max_queue_delay_microseconds": 5000}})
    eng = S.B200LlmPreprocessRequest(ep)
    eng._tokenizer = tokenizer
    return eng, fake


def test_engine_is_registered_with_reference_flags():
    cls = BasePreprocessRequest.get_engine_cls("b200_llm")
    assert cls is S.B200LlmPreprocessRequest
    assert cls.is_preprocess_async and cls.is_process_async and cls.is_postprocess_async     # ps.py:1099-1101
    for name in ("v1_completions", "v1_chat_completions", "v1_models", "version", "tokenize", "detokenize"):
        assert asyncio.iscoroutinefunction(getattr(cls, name))


def test_completions_token_ids_batch_and_usage(monkeypatch):
    eng, fake = _engine(monkeypatch)
    try:
        stats = []
        r = asyncio.run(eng.v1_completions({"request": {"model": "llama", "prompt": [[1, 2, 3], [10]], "max_tokens": 3}}, {}, stats.append))
        assert r["object"] == "text_completion" and r["model"] == "llama"
        assert [c["token_ids"] for c in r["choices"]] == [[6, 7, 8], [10, 11, 12]]
        assert [c["index"] for c in r["choices"]] == [0, 1] and all(c["finish_reason"] == "length" for c in r["choices"])
        assert r["usage"] == {"prompt_tokens": 4, "completion_tokens": 6, "total_tokens": 10}
        assert stats == [{"prompt_tokens": 4, "completion_tokens": 6}]
        assert fake.waves == [(2, 3)]                      # both prompts rode one wave
        one = asyncio.run(eng.v1_completions({"prompt": [4, 4], "max_tokens": 1}, {}, None))
        assert one["choices"][0]["token_ids"] == [8]
    finally:
        eng.unload()
    assert fake.closed


def test_completions_rejects_bad_requests(monkeypatch):
    eng, _ = _engine(monkeypatch)
    try:
        for body in ({"prompt": "hello", "max_tokens": 2},                 # text without a tokenizer
                     {"max_tokens": 2},                                    # no prompt
                     {"prompt": [1, 2], "max_tokens": 0},
                     {"prompt": [1, 200000], "max_tokens": 1},             # id outside the vocabulary
                     {"prompt": list(range(60)), "max_tokens": 10}):       # exceeds max_model_len 64
            with pytest.raises(ValueError):
                asyncio.run(eng.v1_completions(body, {}, None))
        with pytest.raises(ValueError):
            asyncio.run(eng.v1_chat_completions({"messages": [{"role": "user", "content": "hi"}]}, {}, None))
    finally:
        eng.unload()


class _Tok(object):
    def encode(self, s):
        return [ord(c) % 256 for c in s]

    def decode(self, ids):
        return "".join(chr(65 + i % 26) for i in ids)

    def apply_chat_template(self, messages, add_generation_prompt=True, tokenize=True):
        return self.encode("|".join(m["content"] for m in messages))


def test_text_and_chat_with_user_tokenizer(monkeypatch):
    eng, _ = _engine(monkeypatch, tokenizer=_Tok())
    try:
        r = asyncio.run(eng.v1_completions({"prompt": "ab", "max_tokens": 2}, {}, None))
        first = (97 + 98) % 1000
        assert r["choices"][0]["token_ids"] == [first, first + 1] and r["choices"][0]["text"] == _Tok().decode([first, first + 1])
        c = asyncio.run(eng.v1_chat_completions({"messages": [{"role": "user", "content": "a"}], "max_tokens": 1}, {}, None))
        assert c["object"] == "chat.completion" and c["choices"][0]["message"]["role"] == "assistant"
        t = asyncio.run(eng.tokenize({"prompt": "ab"}, {}, None))
        assert t["tokens"] == [97, 98] and t["count"] == 2
        assert asyncio.run(eng.detokenize({"tokens": [0, 1]}, {}, None)) == {"prompt": "AB"}
        assert asyncio.run(eng.v1_models(None, {}, None))["data"][0]["id"] == "llama"
    finally:
        eng.unload()


def test_openai_route_end_to_end(monkeypatch):
    from starlette.testclient import TestClient
    from clearml_serving_b200.main import create_app
    eng, fake = _engine(monkeypatch)
    p = ModelRequestProcessor()
    p._endpoints["llama"] = eng.model_endpoint
    p._engine_processor_lookup["llama"] = eng
    client = TestClient(create_app(p), raise_server_exceptions=False)
    try:
        ok = client.post("/serve/openai/v1/completions", json={"model": "llama", "prompt": [1, 2], "max_tokens": 2})
        assert ok.status_code == 200 and ok.json()["choices"][0]["token_ids"] == [3, 4]
        plain = client.post("/serve/llama", json={"prompt": [7], "max_tokens": 1})
        assert plain.status_code == 200 and plain.json()["choices"][0]["token_ids"] == [7]
        bad = client.post("/serve/openai/v1/completions", json={"model": "llama", "prompt": "text", "max_tokens": 2})
        assert bad.status_code == 422                                       # ValueError -> 422 (main.py:155-161)
        missing = client.post("/serve/openai/v1/completions", json={"model": "nope", "prompt": [1]})
        assert missing.status_code == 404
        wrong_type = client.post("/serve/openai/v1/completions", content=b"x", headers={"Content-Type": "text/plain"})
        assert wrong_type.status_code == 415                                # main.py:208-215
    finally:
        p.shutdown()


def test_concurrent_clients_share_waves(monkeypatch):
    eng, fake = _engine(monkeypatch, max_batch=4)
    fake.latency_s = 0.02

    async def run():
        return await asyncio.gather(*[eng.v1_completions({"prompt": [i], "max_tokens": 2}, {}, None) for i in range(8)])
    try:
        rs = asyncio.run(run())
        assert [r["choices"][0]["token_ids"] for r in rs] == [[i, i + 1] for i in range(8)]
        assert len(fake.waves) <= 4 and sum(n for n, _ in fake.waves) == 8
    finally:
        eng.unload()


def _sse(text):
    import json
    events = [e[len("data: "):] for e in text.split("\n\n") if e.startswith("data: ")]
    assert events[-1] == "[DONE]"
    return [json.loads(e) for e in events[:-1]]


def test_streaming_completions_over_the_openai_route(monkeypatch):
    """"stream": true -> server-sent events (reference: StreamingResponse around the vLLM generator, ps.py:1262-1277):
    tokens arrive in deliveries of the wave's decode loop, per prompt, and add up to the non-streamed answer"""
    from starlette.testclient import TestClient
    from clearml_serving_b200.main import create_app
    eng, fake = _engine(monkeypatch, tokenizer=_Tok())
    p = ModelRequestProcessor()
    p._endpoints["llama"] = eng.model_endpoint
    p._engine_processor_lookup["llama"] = eng
    client = TestClient(create_app(p), raise_server_exceptions=False)
    try:
        body = {"model": "llama", "prompt": [[1, 2, 3], [10]], "max_tokens": 20}
        whole = client.post("/serve/openai/v1/completions", json=body).json()
        r = client.post("/serve/openai/v1/completions", json=dict(body, stream=True))
        assert r.status_code == 200 and r.headers["content-type"].startswith("text/event-stream")
        chunks = _sse(r.text)
        assert all(c["object"] == "text_completion" and len(c["choices"]) == 1 for c in chunks)
        for i in (0, 1):
            mine = [c["choices"][0] for c in chunks if c["choices"][0]["index"] == i]
            assert len(mine) >= 3                                           # prefill token, decode chunks, the tail
            assert sum((m["token_ids"] for m in mine), []) == whole["choices"][i]["token_ids"]
            assert "".join(m["text"] for m in mine) == whole["choices"][i]["text"]
            assert [m["finish_reason"] for m in mine[:-1]] == [None] * (len(mine) - 1) and mine[-1]["finish_reason"] == "length"
        assert len({c["id"] for c in chunks}) == 1 and fake.progress_calls >= 3
        # chat: role on the first delta only, content adds up
        rc = client.post("/serve/openai/v1/chat/completions", json={"model": "llama", "stream": True, "max_tokens": 12,
                                                                    "messages": [{"role": "user", "content": "hi"}]})
        cc = _sse(rc.text)
        assert cc[0]["object"] == "chat.completion.chunk" and cc[0]["choices"][0]["delta"]["role"] == "assistant"
        assert all("role" not in c["choices"][0]["delta"] for c in cc[1:])
        plain = client.post("/serve/openai/v1/chat/completions", json={"model": "llama", "max_tokens": 12,
                                                                       "messages": [{"role": "user", "content": "hi"}]}).json()
        assert "".join(c["choices"][0]["delta"]["content"] for c in cc) == plain["choices"][0]["message"]["content"]
        # a request that cannot run is refused before any event is sent (422, like the non-streamed form)
        bad = client.post("/serve/openai/v1/completions", json={"model": "llama", "prompt": list(range(60)), "max_tokens": 10,
                                                                "stream": True})
        assert bad.status_code == 422
    finally:
        p.shutdown()


def test_streaming_and_plain_requests_share_a_wave(monkeypatch):
    eng, fake = _engine(monkeypatch, max_batch=4)
    fake.latency_s = 0.02
    got = []

    async def run():
        resp = eng._stream([np.array([5], np.int32)], 10, lambda i, toks, text, fin: dict(t=toks, fin=fin))
        plain = asyncio.ensure_future(eng.v1_completions({"prompt": [7], "max_tokens": 4}, {}, None))
        async for ev in resp.body_iterator:
            got.append(ev)
        return await plain
    try:
        r = asyncio.run(run())
        assert r["choices"][0]["token_ids"] == [7, 8, 9, 10]
        toks = sum((e["t"] for e in _sse("".join(got))), [])
        assert toks == list(range(5, 15)) and fake.waves == [(2, 10)]
    finally:
        eng.unload()


# ------------------------------------------------------------------------------------------------
# continuous batching over the paged KV cache
# ------------------------------------------------------------------------------------------------
class FakePagedLlm(object):
    """Host double of LlmEngine.step() that EMULATES the paged cache: every token is stored in pool page
    page_table[slot][pos // 64] at row pos % 64 and the next token is a function of the whole cached sequence READ BACK
    THROUGH THE PAGE TABLE -- a wrong / shared / missing page changes the answer.
    next token = (sum of cached tokens + number of cached tokens) % vocab"""

    def __init__(self, max_batch=4, max_ctx=256, n_pages=None, latency_s=0.0, vocab=1000):
        self.max_batch, self.max_ctx, self.latency_s, self.vocab = max_batch, max_ctx, latency_s, vocab
        self.pages_per_seq = (max_ctx + 63) // 64
        self.n_pages = n_pages or max_batch * self.pages_per_seq
        self.max_prefill_tokens = max_batch * max_ctx
        self.pool = np.full((self.n_pages, 64), -1, np.int64)
        self.table = np.full((max_batch, self.pages_per_seq), -1, np.int64)
        self.calls, self.closed, self.fail_next = [], False, False

        class _Spec(object):
            vocab_size = vocab
        self.spec = _Spec()

    def kv_info(self):
        return self.n_pages, 64, self.pages_per_seq

    def _put(self, slot, pos, tok):
        page = self.table[slot, pos // 64]
        assert 0 <= page < self.n_pages, "slot {} has no page for position {}".format(slot, pos)
        self.pool[page, pos % 64] = tok

    def _next(self, slot, n_cached):
        toks = [self.pool[self.table[slot, p // 64], p % 64] for p in range(n_cached)]
        assert min(toks) >= 0, "read a never-written cache row"
        return int((sum(toks) + n_cached) % self.vocab)

    def step(self, page_updates=(), prefill=None, decode=None):
        if self.fail_next:
            self.fail_next = False
            raise ValueError("injected engine failure")
        time.sleep(self.latency_s)
        for slot, first, pages in page_updates:
            self.table[slot, first:first + len(pages)] = pages
        live = self.table[self.table >= 0]
        first_tokens = new = None
        self.calls.append((len(prefill[0]) if prefill else 0, len(decode[0]) if decode else 0, decode[3] if decode else 0))
        if prefill is not None:
            prompts, slots = prefill
            assert len(set(slots)) == len(slots)
            first_tokens = []
            for p, sl in zip(prompts, slots):
                for i, t in enumerate(p):
                    self._put(sl, i, int(t))
                first_tokens.append(self._next(sl, len(p)))
            first_tokens = np.asarray(first_tokens, np.int32)
        if decode is not None:
            slots, ctx, last, n = decode
            new = np.zeros((len(slots), n), np.int32)
            for r, (sl, c, t) in enumerate(zip(slots, ctx, last)):
                for k in range(n):
                    self._put(sl, c + k, int(t))
                    t = self._next(sl, c + k + 1)
                    new[r, k] = t
        return first_tokens, new

    def close(self):
        self.closed = True


def _expected(prompt, n, vocab=1000):
    seq, out = [int(t) for t in prompt], []
    for _ in range(n):
        t = (sum(seq) + len(seq)) % vocab
        out.append(t)
        seq.append(t)
    return out


def test_continuous_batcher_joins_and_leaves_at_step_granularity():
    eng = FakePagedLlm(max_batch=4, max_ctx=256, latency_s=0.002)
    b = S.ContinuousBatcher(eng, max_batch=4, max_ctx=256, chunk=3)
    rng = np.random.default_rng(0)
    try:
        reqs = [(rng.integers(0, 1000, int(rng.integers(1, 150))), int(rng.integers(1, 40))) for _ in range(24)]
        futs = []
        for i, (p, n) in enumerate(reqs):
            futs.append(b.submit(p, n))
            if i % 5 == 4:
                time.sleep(0.01)                          # later arrivals find a running batch
        for (p, n), f in zip(reqs, futs):
            r = f.result(timeout=30)
            assert r.tolist() == _expected(p, n) and r.finish_reason == "length"
        st = b.stats
        assert st["requests"] == 24 and st["joined_running"] > 0             # newcomers were prefilled next to running sequences
        assert st["max_rows"] <= 4 and st["decode_steps"] > 0
        assert max(c[0] + c[1] for c in eng.calls) <= 4                      # never more sequences than KV slots in one iteration
        assert sorted(b._free_slots) == [0, 1, 2, 3] and sorted(b._free_pages) == list(range(eng.n_pages)) and b._reserved == 0
        # a sequence never decodes past its own max_tokens: the chunk is cut at the nearest finish
        assert all(c[2] <= 3 for c in eng.calls)
    finally:
        b.close()


def test_continuous_batcher_small_page_pool_and_stop_tokens():
    eng = FakePagedLlm(max_batch=4, max_ctx=256, n_pages=5)                  # 5 pages = 320 tokens for everybody
    b = S.ContinuousBatcher(eng, max_batch=4, max_ctx=256, chunk=4)
    try:
        with pytest.raises(ValueError, match="max_model_len"):
            b.submit(np.arange(250), 10)
        prompts = [np.full(100, i + 1) for i in range(4)]                    # each needs ceil((100 + 28) / 64) = 2 pages
        futs = [b.submit(p, 28) for p in prompts]
        for p, f in zip(prompts, futs):
            assert f.result(timeout=30).tolist() == _expected(p, 28)
        assert b.stats["pages_peak"] <= 5 and b.stats["max_rows"] <= 2       # the pool, not the slots, bounded the batch
        assert sorted(b._free_pages) == list(range(5)) and b._reserved == 0
        # stop tokens: generation ends with the first stop token, which is part of the answer
        p = np.array([3, 4, 5])
        want = _expected(p, 20)
        r = b.submit(p, 20, None, [want[6]]).result(timeout=30)
        k = want.index(want[6])
        assert r.tolist() == want[:k + 1] and r.finish_reason == "stop" and b.stats["finished_stop"] == 1
        # an engine failure fails the sequences of that iteration only; the endpoint keeps serving
        eng.fail_next = True
        with pytest.raises(ValueError, match="injected"):
            b.submit(np.array([1, 2]), 3).result(timeout=30)
        assert b.submit(np.array([1, 2]), 3).result(timeout=30).tolist() == _expected([1, 2], 3)
        assert sorted(b._free_slots) == [0, 1, 2, 3] and sorted(b._free_pages) == list(range(5))
    finally:
        b.close()
    with pytest.raises(RuntimeError):
        b.submit(np.array([1]), 1)


def test_openai_surface_on_the_continuous_scheduler(monkeypatch):
    fake = FakePagedLlm(max_batch=4, max_ctx=128)
    eng, _ = _engine(monkeypatch, tokenizer=_Tok(), scheduler="continuous", fake=fake, extra={"eos_token_id": 777})
    assert isinstance(eng._batcher, S.ContinuousBatcher)
    try:
        async def run():
            return await asyncio.gather(*[eng.v1_completions({"prompt": [i + 1, 2], "max_tokens": 5 + i}, {}, None) for i in range(6)])
        rs = asyncio.run(run())
        for i, r in enumerate(rs):
            assert r["choices"][0]["token_ids"] == _expected([i + 1, 2], 5 + i) and r["choices"][0]["finish_reason"] == "length"
        # EOS ends the generation ("stop"); ignore_eos generates on; stop_token_ids add to it
        want = _expected([9, 9], 12)
        eng._eos_ids = (want[3],)
        r = asyncio.run(eng.v1_completions({"prompt": [9, 9], "max_tokens": 12}, {}, None))
        assert r["choices"][0]["token_ids"] == want[:4] and r["choices"][0]["finish_reason"] == "stop"
        assert r["usage"]["completion_tokens"] == 4
        r = asyncio.run(eng.v1_completions({"prompt": [9, 9], "max_tokens": 12, "ignore_eos": True}, {}, None))
        assert r["choices"][0]["token_ids"] == want and r["choices"][0]["finish_reason"] == "length"
        r = asyncio.run(eng.v1_completions({"prompt": [9, 9], "max_tokens": 12, "ignore_eos": True, "stop_token_ids": [want[1]]}, {}, None))
        assert r["choices"][0]["token_ids"] == want[:2]
        # streamed: the last event carries the real finish reason
        async def stream():
            resp = await eng.v1_completions({"prompt": [9, 9], "max_tokens": 12, "stream": True}, {}, None)
            return "".join([ev async for ev in resp.body_iterator])
        chunks = _sse(asyncio.run(stream()))
        assert sum((c["choices"][0]["token_ids"] for c in chunks), []) == want[:4] and chunks[-1]["choices"][0]["finish_reason"] == "stop"
        # what a greedy engine cannot honour is refused, not silently ignored
        for bad in ({"temperature": 0.7}, {"n": 2}, {"top_p": 0.9}, {"stop": ["\n"]}, {"max_tokens": -3}):
            with pytest.raises(ValueError):
                asyncio.run(eng.v1_completions(dict({"prompt": [1, 2], "max_tokens": 2}, **bad), {}, None))
        with pytest.raises(ValueError):
            asyncio.run(eng.v1_chat_completions({"messages": [{"role": "user", "content": "a"}], "max_tokens": -1}, {}, None))
        ok = asyncio.run(eng.v1_completions({"prompt": [1, 2], "max_tokens": 2, "temperature": 0, "top_p": 1.0, "n": 1}, {}, None))
        assert ok["choices"][0]["token_ids"] == _expected([1, 2], 2)
    finally:
        eng.unload()
    assert fake.closed


def test_wave_scheduler_does_not_let_one_request_fail_its_wave_mates():
    """regression: a wave generates max(max_tokens) for every member -- A (long prompt, few tokens) and B (short
    prompt, many tokens) are each valid alone and used to get a 422 together"""
    eng = FakeLlm(max_batch=4, max_ctx=64)

    def checked_generate(prompts, n, **kw):
        assert all(len(p) + n <= 64 for p in prompts), "wave violates the context limit"
        return FakeLlm.generate(eng, prompts, n, **kw)
    eng.generate = checked_generate
    b = S.WaveBatcher(eng, max_batch=4, max_queue_delay_us=30000)
    try:
        fa = b.submit(np.arange(60), 2)
        fb = b.submit(np.array([1, 2]), 50)
        assert len(fa.result(timeout=10)) == 2 and len(fb.result(timeout=10)) == 50
        assert len(eng.waves) == 2
    finally:
        b.close()


def test_tensor_parallel_leader_refuses_before_announcing():
    """a doomed call must not reach the follower: it would raise inside follower_loop and the leader's next broadcast would
    hang forever"""
    from clearml_serving_b200 import llm as L

    class _Inner(object):
        max_ctx = 32
        _check = L.LlmEngine._check

        def generate(self, *a, **k):
            raise AssertionError("must not be reached")
    leader = S.TensorParallelLeader(_Inner())
    announced = []
    leader._announce = announced.append
    with pytest.raises(ValueError, match="max_ctx"):
        leader.generate([list(range(30))], 10)
    assert announced == []
