"""Engine metrics export: the lines the reference's Triton sidecar parses from tritonserver's
:8002/metrics (clearml_serving/engines/triton/triton_helper.py:20,:45-89) come from the b200 batcher counters, and a
sampled request reports its batch figures on the reference's statistics channel (model_request_processor.py:1341-1367)."""
import asyncio
import re

import numpy as np

from clearml_serving_b200 import ModelEndpoint, metrics
from clearml_serving_b200.model_request_processor import ModelRequestProcessor
from clearml_serving_b200.scheduler import BATCH_ROWS_BUCKETS
from tests.fakes import FakeModel, make_fake_engine

# the reference's line format, verbatim (triton_helper.py:20)
REF_METRIC_LINE = re.compile(r"(\w+){(gpu_uuid=\"[\w\W]*\",)?model=\"(\w+)\",\s*version=\"(\d+)\"}\s*([0-9.]*)")


def _processor():
    p = ModelRequestProcessor()
    for url, ver in (("sum", ""), ("sum/2", "2")):
        ep = ModelEndpoint(engine_type="b200", serving_url="sum", version=ver)
        p._endpoints[url] = ep
        p._engine_processor_lookup[url] = make_fake_engine(ep, FakeModel(n_features=4))
    return p


def test_metrics_endpoint_speaks_the_triton_sidecar_format():
    from starlette.testclient import TestClient
    from clearml_serving_b200.main import create_app
    p = _processor()
    client = TestClient(create_app(p), raise_server_exceptions=False)
    try:
        for i in range(5):
            assert client.post("/serve/sum", content=__import__("clearml_serving_b200").wire.encode_tensors([np.array([[1, 2, 3, i]], np.float32)]), headers={"Content-Type": "application/x-b200-tensors"}).status_code == 200
        eng = p._engine_processor_lookup["sum"]
        for i in range(7):
            eng.process_sync([[1.0, 2.0, 3.0, float(i)]])
        eng.process_sync(np.ones((3, 4), np.float32))
        r = client.get("/metrics")
        assert r.status_code == 200 and r.headers["content-type"].startswith("text/plain")
        parsed = {}
        for line in r.text.split("\n"):
            line = line.strip()
            if not line or line.startswith("#"):
                continue
            m = REF_METRIC_LINE.match(line)
            if m and "le=" not in line:
                metric, _gpu, model, version, value = m.groups()
                parsed[(metric, model, version)] = float(value)
        st = eng.engine_stats()
        assert parsed[("nv_inference_request_success", "sum", "1")] == st["requests"] >= 8
        assert parsed[("nv_inference_count", "sum", "1")] == st["rows"] >= 10
        assert parsed[("nv_inference_exec_count", "sum", "1")] == st["batches"]
        assert parsed[("nv_inference_request_failure", "sum", "1")] == 0
        assert parsed[("nv_inference_queue_duration_us", "sum", "1")] >= 0
        assert parsed[("nv_inference_compute_infer_duration_us", "sum", "1")] > 0
        assert parsed[("b200_input_bytes", "sum", "1")] == st["in_bytes"] >= 10 * 16
        assert ("nv_inference_count", "sum_2", "2") in parsed          # model name rule: "{url}_{version}" (ps.py:375-379)
        # histogram: cumulative buckets, +Inf == count == batches
        buckets = re.findall(r'b200_batch_rows_bucket\{model="sum",version="1",le="([^"]+)"\} (\d+)', r.text)
        assert [b for b, _ in buckets] == [str(b) for b in BATCH_ROWS_BUCKETS] + ["+Inf"]
        counts = [int(c) for _, c in buckets]
        assert counts == sorted(counts) and counts[-1] == st["batches"] == sum(st["batch_rows_hist"])
    finally:
        p.shutdown()


def test_sampled_request_reports_its_batch_figures():
    ep = ModelEndpoint(engine_type="b200", serving_url="m")
    eng = make_fake_engine(ep, FakeModel(n_features=2), latency_s=0.002)
    try:
        got = {}

        async def one():
            return await eng.process([[1.0, 2.0], [3.0, 4.0]], {}, got.update)
        out = asyncio.run(one())
        assert np.array_equal(out, [3.0, 7.0])
        assert got["_b200_batch_rows"] == 2 and got["_b200_queue_us"] >= 0 and got["_b200_exec_us"] >= 2000
        # an unsampled request takes the plain path and reports nothing
        got.clear()
        asyncio.run(eng.process([[1.0, 1.0]], {}, None))
        assert got == {}
    finally:
        eng.unload()


def test_failed_batches_are_counted():
    from clearml_serving_b200.scheduler import BatchPolicy, DynamicBatcher
    from tests.fakes import FakeStream
    model = FakeModel(n_features=2)
    pol = BatchPolicy(max_batch_size=4, max_queue_delay_us=0)
    b = DynamicBatcher(model, pol, name="f", stream=FakeStream(model, 4, fail_on=1))
    try:
        f = b.submit([np.ones((1, 2), np.float32)], 1)
        try:
            f.result(timeout=5)
            raise AssertionError("the injected failure did not propagate")
        except ValueError:
            pass
        assert b.submit([np.ones((1, 2), np.float32)], 1).result(timeout=5)[0][0] == 2.0
        st = b.snapshot_stats()
        assert st["failed_requests"] == 1 and st["requests"] - st["failed_requests"] == 1
        text = metrics.render({"f": ("1", st)})
        assert 'nv_inference_request_failure{model="f",version="1"} 1' in text
        assert 'nv_inference_request_success{model="f",version="1"} 1' in text
    finally:
        b.shutdown()
