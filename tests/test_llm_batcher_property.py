"""Property test of the continuous-batching scheduler (llm_service.ContinuousBatcher) over random request streams, pool
sizes and chunk lengths, on the page-table-emulating engine double of tests/test_llm_service.py: whatever the mix,
 * every request gets exactly the tokens it would get alone, cut after its first stop token;
 * an iteration never carries more sequences than KV slots, the pool never hands out more pages than it has;
 * afterwards every slot and page is back and nothing stays reserved."""
import time

import numpy as np
import pytest

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import HealthCheck, given, settings, strategies as st  # noqa: E402

from clearml_serving_b200 import llm_service as S  # noqa: E402
from tests.test_llm_service import FakePagedLlm, _expected  # noqa: E402

request = st.tuples(st.integers(1, 120), st.integers(1, 40), st.integers(0, 3), st.booleans())


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(max_batch=st.integers(1, 4), n_pages=st.integers(3, 12), chunk=st.integers(1, 5),
       reqs=st.lists(request, min_size=1, max_size=14), seed=st.integers(0, 10 ** 6))
def test_random_streams_keep_the_invariants(max_batch, n_pages, chunk, reqs, seed):
    rng = np.random.default_rng(seed)
    eng = FakePagedLlm(max_batch=max_batch, max_ctx=192, n_pages=n_pages)
    b = S.ContinuousBatcher(eng, max_batch=max_batch, max_ctx=192, chunk=chunk)
    try:
        work = []
        for plen, gen, stop_at, pause in reqs:
            prompt = rng.integers(0, 1000, plen)
            want = _expected(prompt, gen)
            stops = [want[min(stop_at + 2, gen - 1)]] if stop_at else []      # a token that WILL be generated, or none
            pages = -(-(plen + gen) // 64)
            if plen + gen > 192 or pages > n_pages:
                with pytest.raises(ValueError):
                    b.submit(prompt, gen, None, stops)
                continue
            work.append((want, stops, b.submit(prompt, gen, None, stops)))
            if pause:
                time.sleep(0.002)                                              # some arrivals find a running batch
        for want, stops, fut in work:
            r = fut.result(timeout=60)
            if stops:
                k = min(want.index(s) for s in stops if s in want)
                assert r.tolist() == want[:k + 1] and r.finish_reason == "stop"
            else:
                assert r.tolist() == want and r.finish_reason == "length"
        assert all(c[0] + c[1] <= max_batch for c in eng.calls)
        assert b.stats["pages_peak"] <= n_pages and b.stats["max_rows"] <= max_batch
        assert sorted(b._free_slots) == list(range(max_batch)) and sorted(b._free_pages) == list(range(n_pages)) and b._reserved == 0
    finally:
        b.close()
