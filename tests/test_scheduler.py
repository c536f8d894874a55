"""Dynamic batcher (scheduler.py) semantics on CPU with a fake stream (SURVEY.md 5.9 rules)."""
import asyncio
import time

import numpy as np
import pytest

from clearml_serving_b200.scheduler import BatchPolicy, DynamicBatcher
from tests.fakes import FakeModel, FakeStream


def test_policy_from_cli_dotted_keys():
    # the exact --aux-config of examples/huggingface/readme.md:113
    aux = {"platform": "onnxruntime_onnx", "default_model_filename": "model.bin",
           "dynamic_batching.preferred_batch_size": "[1,2,4,8,16,32,64]",
           "dynamic_batching.max_queue_delay_microseconds": 5000, "max_batch_size": 64}
    p = BatchPolicy.from_auxiliary_cfg(aux)
    assert (p.max_batch_size, p.max_queue_delay_us) == (64, 5000)
    assert p.preferred_batch_size == [1, 2, 4, 8, 16, 32, 64]


def test_policy_from_nested_dict_and_pbtxt():
    p = BatchPolicy.from_auxiliary_cfg({"max_batch_size": 8, "dynamic_batching": {"max_queue_delay_microseconds": "100"}})
    assert (p.max_batch_size, p.max_queue_delay_us) == (8, 100)
    text = 'platform: "pytorch_libtorch"\nmax_batch_size: 128\ndynamic_batching { preferred_batch_size: [ 4, 8 ]\n max_queue_delay_microseconds: 250 }\n'
    p = BatchPolicy.from_auxiliary_cfg(text)
    assert (p.max_batch_size, p.max_queue_delay_us, p.preferred_batch_size) == (128, 250, [4, 8])
    p = BatchPolicy.from_auxiliary_cfg(None)
    assert (p.max_batch_size, p.max_queue_delay_us) == (64, 0)
    assert BatchPolicy.from_auxiliary_cfg({"max_batch_size": 0}).max_batch_size == 1


def _mk(policy, **kw):
    model = FakeModel(n_features=3)
    stream = FakeStream(model, policy.max_batch_size, n_slots=policy.n_slots, **kw)
    return DynamicBatcher(model, policy, name="t", stream=stream), stream


def test_results_are_per_request_and_fifo():
    b, s = _mk(BatchPolicy(max_batch_size=8, max_queue_delay_us=20000))
    try:
        xs = [np.full((1 + i % 3, 3), i, np.float32) for i in range(20)]
        futs = [b.submit([x], x.shape[0]) for x in xs]
        for i, (x, f) in enumerate(zip(xs, futs)):
            out = f.result(timeout=5)[0]
            assert out.shape == (x.shape[0],) and np.all(out == 3 * i)
        assert max(s.batches) <= 8 and sum(s.batches) == sum(x.shape[0] for x in xs)
        assert len(s.batches) < 20  # batching happened
    finally:
        b.shutdown()


def test_delay_zero_dispatches_immediately_and_delay_waits():
    b, s = _mk(BatchPolicy(max_batch_size=64, max_queue_delay_us=0))
    try:
        t0 = time.perf_counter()
        b.submit([np.ones((1, 3), np.float32)], 1).result(timeout=5)
        assert time.perf_counter() - t0 < 0.2
    finally:
        b.shutdown()
    b, s = _mk(BatchPolicy(max_batch_size=64, max_queue_delay_us=150000))
    try:
        t0 = time.perf_counter()
        f1 = b.submit([np.ones((1, 3), np.float32)], 1)
        time.sleep(0.02)
        f2 = b.submit([np.ones((2, 3), np.float32)], 2)
        f1.result(timeout=5), f2.result(timeout=5)
        dt = time.perf_counter() - t0
        assert 0.1 < dt < 1.0          # waited for the queue delay of the OLDEST request
        assert s.batches == [3]        # ... and both requests went out together
    finally:
        b.shutdown()


def test_full_batch_goes_without_waiting():
    b, s = _mk(BatchPolicy(max_batch_size=4, max_queue_delay_us=2000000))
    try:
        t0 = time.perf_counter()
        futs = [b.submit([np.ones((1, 3), np.float32)], 1) for _ in range(4)]
        [f.result(timeout=5) for f in futs]
        assert time.perf_counter() - t0 < 1.0
        assert s.batches[0] == 4
    finally:
        b.shutdown()


def test_preferred_batch_size_is_used():
    b, s = _mk(BatchPolicy(max_batch_size=16, max_queue_delay_us=100000, preferred_batch_size=[4, 8]))
    try:
        futs = [b.submit([np.ones((1, 3), np.float32)], 1) for _ in range(6)]
        [f.result(timeout=5) for f in futs]
        assert s.batches[0] == 4 and sum(s.batches) == 6
    finally:
        b.shutdown()


def test_oversized_request_rejected_and_failure_isolated():
    b, s = _mk(BatchPolicy(max_batch_size=4), fail_on=1)
    try:
        with pytest.raises(ValueError, match="max_batch_size"):
            b.submit([np.ones((5, 3), np.float32)], 5)
        f1 = b.submit([np.ones((1, 3), np.float32)], 1)
        with pytest.raises(ValueError, match="injected"):
            f1.result(timeout=5)
        f2 = b.submit([np.ones((1, 3), np.float32)], 1)   # the next batch is unaffected
        assert f2.result(timeout=5)[0][0] == 3
    finally:
        b.shutdown()


def test_asyncio_path_many_concurrent_requests():
    b, s = _mk(BatchPolicy(max_batch_size=32, max_queue_delay_us=1000, n_slots=2), latency_s=0.002)

    async def main():
        async def one(i):
            out = await b.submit_async([np.full((1, 3), i, np.float32)], 1)
            return float(out[0][0])
        return await asyncio.gather(*[one(i) for i in range(300)])

    try:
        res = asyncio.run(main())
        assert res == [3.0 * i for i in range(300)]
        st = b.snapshot_stats()
        assert st["requests"] == 300 and st["rows"] == 300 and st["mean_batch_rows"] > 1
    finally:
        b.shutdown()


def test_random_request_streams_keep_the_batcher_invariants():
    """property test over request streams (row counts, arrival gaps, policies): every request gets exactly its own rows
    back, no batch exceeds max_batch_size, rows are conserved, and requests leave in arrival order (FIFO batches)"""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=25, deadline=None)
    @given(st.lists(st.tuples(st.integers(1, 6), st.integers(0, 3)), min_size=1, max_size=40),
           st.sampled_from([(8, 0, ()), (8, 500, ()), (16, 300, (4, 8)), (6, 0, (2, 4))]))
    def run(reqs, pol):
        max_bs, delay, preferred = pol
        model = FakeModel(n_features=3)
        stream = FakeStream(model, max_bs, n_slots=2)
        b = DynamicBatcher(model, BatchPolicy(max_batch_size=max_bs, max_queue_delay_us=delay, preferred_batch_size=list(preferred),
                                              n_slots=2), name="prop", stream=stream)
        try:
            futs, wants, tag = [], [], 0
            for rows, gap in reqs:
                x = np.full((rows, 3), 0.0, np.float32)
                x[:, 0] = np.arange(tag, tag + rows)            # a unique value per row across the whole stream
                tag += rows
                futs.append(b.submit([x], rows))
                wants.append(x.sum(axis=1))
                if gap:
                    time.sleep(gap * 1e-4)
            outs = [f.result(timeout=10) for f in futs]
            for o, w in zip(outs, wants):
                assert len(o) == 1 and np.array_equal(o[0], w)
            assert all(0 < n <= max_bs for n in stream.batches)
            assert sum(stream.batches) == tag
            st_ = b.snapshot_stats()
            assert st_["requests"] == len(reqs) and st_["rows"] == tag and st_["batches"] == len(stream.batches)
            assert sum(st_["batch_rows_hist"]) == st_["batches"]
        finally:
            b.shutdown()
    run()


def test_failed_batches_give_their_staging_slot_back():
    """regression: a submit / collate failure used to leave the slot ACQUIRED; after n_slots failures the
    dispatcher spun in _acquire_slot forever and every later request of the endpoint hung"""
    b, s = _mk(BatchPolicy(max_batch_size=4, n_slots=2), fail_on={1, 2, 3})
    try:
        for _ in range(3):
            f = b.submit([np.ones((1, 3), np.float32)], 1)
            with pytest.raises(ValueError, match="injected"):
                f.result(timeout=5)
        assert sorted(s.free) == [0, 1]
        ok = b.submit([np.ones((2, 3), np.float32)], 2)
        assert np.array_equal(ok.result(timeout=5)[0], [3.0, 3.0])
        assert b.snapshot_stats()["failed_requests"] == 3
    finally:
        b.shutdown()


def test_expired_requests_do_not_form_an_empty_batch():
    """every queued request older than request_timeout_s: all of them fail with the timeout error, nothing is
    dispatched, and the endpoint keeps serving (the head-alone path honours the deadline too)"""
    model = FakeModel(n_features=3)
    stream = FakeStream(model, 4, n_slots=1, latency_s=0.15)
    b = DynamicBatcher(model, BatchPolicy(max_batch_size=4, n_slots=1, preferred_batch_size=[4]), name="exp", stream=stream,
                       request_timeout_s=0.05)
    try:
        first = b.submit([np.ones((4, 3), np.float32)], 4)          # occupies the only slot for 150 ms
        time.sleep(0.01)
        late = [b.submit([np.ones((3, 3), np.float32)], 3) for _ in range(3)]   # 3 rows: each goes "head alone"
        assert first.result(timeout=5)[0].shape == (4,)
        for f in late:
            with pytest.raises(ValueError, match="timed out"):
                f.result(timeout=5)
        fresh = b.submit([np.ones((1, 3), np.float32)], 1)
        assert fresh.result(timeout=5)[0][0] == 3
        assert all(n > 0 for n in stream.batches)
    finally:
        b.shutdown()
