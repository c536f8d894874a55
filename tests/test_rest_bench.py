"""bench.py's REST-level load generator (level L1: uvicorn + an aiohttp closed-loop client in one process) on
CPU against the host double: replies are checked one by one, failures are counted, the server goes away afterwards."""
import json
import os
import sys

import numpy as np

from clearml_serving_b200 import ModelEndpoint, wire
from clearml_serving_b200.model_request_processor import ModelRequestProcessor
from tests.fakes import FakeModel, make_fake_engine

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def _app():
    """a fresh processor per run: the app's shutdown event (uvicorn stopping) tears the processor's engines down"""
    from clearml_serving_b200.main import create_app
    p = ModelRequestProcessor()
    ep = ModelEndpoint(engine_type="b200", serving_url="sum")
    p._endpoints["sum"] = ep
    p._engine_processor_lookup["sum"] = make_fake_engine(ep, FakeModel(n_features=4))
    return p, create_app(p)


def test_rest_load_generator_counts_and_checks():
    X = np.arange(40, dtype=np.float32).reshape(10, 4)
    bodies = [wire.encode_tensors([X[i:i + 1]]) for i in range(10)]
    hdr = {"Content-Type": wire.MEDIA_TYPE}
    seen = []

    def check(payload, i):
        seen.append(i)
        return bool(wire.decode_tensors(payload)[0][0] == X[i].sum())
    p, app = _app()
    try:
        r = bench.rest_load(app, "/serve/sum", bodies, hdr, seconds=0.5, concurrency=8, check=check)
    finally:
        p.shutdown()
    assert r["failed"] == 0 and r["mismatched"] == 0 and r["completed"] > 20 and r["req_s"] > 0 and r["p99_us"] >= r["p50_us"] > 0
    assert len(seen) >= r["completed"] and set(seen) == set(range(10))
    p, app = _app()
    try:
        off = bench.rest_load(app, "/serve/sum", bodies, hdr, seconds=0.2, concurrency=2, check=lambda payload, i: False)
    finally:
        p.shutdown()
    assert off["failed"] == 0 and off["mismatched"] == off["completed"] > 0     # replies the checker rejects are counted
    p, app = _app()
    try:
        bad = bench.rest_load(app, "/serve/nope", [json.dumps({"x": 1}).encode()], {"Content-Type": "application/json"},
                              seconds=0.2, concurrency=2)
    finally:
        p.shutdown()
    assert bad["failed"] == bad["completed"] > 0                                 # 404s are counted, not raised
