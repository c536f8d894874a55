"""Engine plugin layer on CPU: registry, user-code loading, marshalling rules (vs the reference's
Triton client recorded in tests/golden/triton_marshal.json), router semantics, REST contract
(vs tests/golden/rest_contract.json recorded from the reference FastAPI app)."""
import asyncio
import gzip
import json
import os
import textwrap
import time

import numpy as np
import pytest

from clearml_serving_b200 import BasePreprocessRequest, ModelEndpoint
from clearml_serving_b200.model_request_processor import (
    EndpointBackendEngineException, EndpointNotFoundException, ModelRequestProcessor)
from tests.fakes import FakeModel, make_fake_engine


def test_registry_and_flags():
    assert BasePreprocessRequest.validate_engine_type("b200")
    assert not BasePreprocessRequest.validate_engine_type("nope")
    cls = BasePreprocessRequest.get_engine_cls("b200")
    # same async profile as the Triton engine (preprocess_service.py:289-291)
    assert (cls.is_preprocess_async, cls.is_process_async, cls.is_postprocess_async) == (False, True, False)
    with pytest.raises(TypeError, match="not supported engine type"):
        ModelEndpoint(engine_type="nope", serving_url="x")

    @BasePreprocessRequest.register_engine("unit_test_engine", modules=["json", "definitely_not_a_module"])
    class E(BasePreprocessRequest):
        pass
    assert BasePreprocessRequest.get_engine_cls("unit_test_engine") is E
    BasePreprocessRequest.load_modules()  # missing modules are ignored silently


def test_endpoint_normalisation():
    ep = ModelEndpoint(engine_type="b200", serving_url="m", input_size=[-1, 4], input_type="float32",
                       input_name="x", output_size=[[1]], output_type=["float32"], output_name=["y"])
    assert ep.input_size == [[-1, 4]] and ep.input_type == ["float32"] and ep.input_name == ["x"]
    assert ep.as_dict(remove_null_entries=True)["serving_url"] == "m"
    with pytest.raises(TypeError):
        ModelEndpoint(engine_type="b200", serving_url="m", input_type=["not_a_dtype"])


def test_user_code_loading_and_unload(tmp_path):
    code = textwrap.dedent('''
        EVENTS = []
        class Preprocess(object):
            def load(self, local_file_name):
                EVENTS.append(("load", local_file_name)); return "the-model"
            def unload(self):
                EVENTS.append(("unload",))
            def preprocess(self, body, state, collect_custom_statistics_fn=None):
                state["k"] = body["v"] * 2; return body["v"]
            def process(self, data, state, collect_custom_statistics_fn=None):
                return data + 1
            def postprocess(self, data, state, collect_custom_statistics_fn=None):
                return dict(y=data, k=state["k"], has_send=callable(self.send_request), ep=self.model_endpoint.serving_url)
    ''')
    path = tmp_path / "preprocess.py"
    path.write_text(code)
    ep = ModelEndpoint(engine_type="custom", serving_url="user", preprocess_artifact=str(path))
    eng = BasePreprocessRequest.get_engine_cls("custom")(model_endpoint=ep, task=None)
    assert eng._model == "the-model"
    st = {}
    out = eng.postprocess(eng.process(eng.preprocess({"v": 5}, st), st), st)
    assert out == dict(y=6, k=10, has_send=True, ep="user")
    events = __import__("sys").modules[type(eng._preprocess).__mro__[1].__module__].EVENTS \
        if type(eng._preprocess).__mro__[1].__module__ in __import__("sys").modules else None
    user_mod_events = eng._preprocess.load.__func__.__globals__["EVENTS"]
    assert user_mod_events[0][0] == "load"
    del eng
    import gc
    gc.collect()
    assert ("unload",) in user_mod_events   # dropping the engine released the user's model

    with pytest.raises(ValueError, match="could not find preprocessing artifact"):
        BasePreprocessRequest.get_engine_cls("custom")(
            model_endpoint=ModelEndpoint(engine_type="custom", serving_url="u", preprocess_artifact="missing_artifact"),
            task=None)


# ---------------------------------------------------------------- marshalling vs the reference client
def _golden_cases(golden_dir):
    with open(os.path.join(golden_dir, "triton_marshal.json")) as f:
        return json.load(f)


def test_marshalling_matches_reference_triton_client(golden_dir):
    """For each recorded case: the arrays our engine hands to the batcher have the dtype/shape the
    reference put on the wire; unsupported dtypes raise the same message."""
    for case in _golden_cases(golden_dir):
        io = case["io"]
        ep = ModelEndpoint(engine_type="b200", serving_url=case["name"], **io)
        n_in = len(io["input_name"])

        class M(FakeModel):
            pass
        m = M()
        m.n_inputs = n_in
        if "raises" in case:
            m.in_dtypes = [np.dtype(np.float32)] * n_in
            m.in_row_elems = [2] * n_in
            eng = make_fake_engine(ep, m)
            try:
                with pytest.raises(ValueError) as ei:
                    eng._marshal(case["data"])
                assert str(ei.value) == case["message"]
            finally:
                eng.unload()
            continue
        wire = case["wire_inputs"]
        np_of = {"FP32": np.float32, "FP64": np.float64, "INT32": np.int32, "INT64": np.int64, "UINT8": np.uint8}
        m.in_dtypes = [np.dtype(np_of[w["datatype"]]) for w in wire]
        # rows = leading dim when the wire tensor is >= 2-D (Triton batch dim)
        shapes = [w["shape"] for w in wire]
        if any(len(s) < 2 for s in shapes):
            continue  # 1-D requests carry no batch dim: rejected by the b200 engine (see test below)
        m.in_row_elems = [int(np.prod(s[1:])) for s in shapes]
        eng = make_fake_engine(ep, m)
        try:
            arrays, rows = eng._marshal(case["data"])
            assert rows == shapes[0][0]
            for a, w, dt in zip(arrays, wire, m.in_dtypes):
                assert a.dtype == dt and a.shape[0] == w["shape"][0] and a.size == int(np.prod(w["shape"]))
        finally:
            eng.unload()


def test_one_dimensional_request_is_rejected_like_sklearn():
    ep = ModelEndpoint(engine_type="b200", serving_url="m")
    eng = make_fake_engine(ep, FakeModel(n_features=3))
    try:
        with pytest.raises(ValueError, match="Expected 2D array"):
            eng._marshal([1.0, 2.0, 3.0])
        with pytest.raises(ValueError, match="features per row"):
            eng._marshal([[1.0, 2.0]])
        arrays, rows = eng._marshal([[1, None, 3]])   # None -> NaN, like DMatrix / np.array(float32)
        assert rows == 1 and np.isnan(arrays[0][0, 1])
    finally:
        eng.unload()


def test_output_dtype_clamp_and_single_output_unwrapped():
    ep = ModelEndpoint(engine_type="b200", serving_url="m", input_size=[[3]], input_type=["float32"], input_name=["x"],
                       output_size=[[1]], output_type=["float64"], output_name=["y"])
    eng = make_fake_engine(ep, FakeModel(n_features=3))
    try:
        out = eng.process_sync([[1, 2, 3], [4, 5, 6]])
        assert isinstance(out, np.ndarray) and out.dtype == np.float64 and out.tolist() == [6.0, 15.0]
        assert out.flags["OWNDATA"] or out.base is not None
    finally:
        eng.unload()


# ---------------------------------------------------------------- router
def _router_with_fake(latency_s=0.0):
    p = ModelRequestProcessor()
    ep = ModelEndpoint(engine_type="b200", serving_url="sum")
    p._endpoints["sum"] = ep
    p._engine_processor_lookup["sum"] = make_fake_engine(ep, FakeModel(n_features=2), latency_s=latency_s)
    return p


def test_router_process_request_and_errors():
    p = _router_with_fake()

    async def main():
        r = await p.process_request("sum", None, [[1, 2]], "process")
        assert r.tolist() == [3.0]
        with pytest.raises(EndpointNotFoundException, match="Model inference endpoint 'nope/3' not found"):
            await p.process_request("nope", "3", {}, "process")
        assert p._request_processing_state.value() == 0
    try:
        asyncio.run(main())
    finally:
        p.shutdown()


def test_router_url_normalisation_and_canary():
    f = ModelRequestProcessor._normalize_endpoint_url
    assert f("a", None) == "a" and f("a/", "") == "a" and f("a", "2") == "a/2" and f("a/", "2") == "a/2"
    p = ModelRequestProcessor()
    for v in ("1", "2", "10"):
        ep = ModelEndpoint(engine_type="b200", serving_url="m/{}".format(v), version=v)
        p._endpoints["m/{}".format(v)] = ep
    p.add_canary_endpoint(dict(endpoint="m", weights=[0.75, 0.25], load_endpoint_prefix="m/"))
    assert p._canary_route["m"]["endpoints"] == ["m/10", "m/2"]      # newest versions first
    assert p._canary_route["m"]["weights"] == [0.75, 0.25]
    draws = [p._process_canary("m") for _ in range(400)]
    assert set(draws) == {"m/10", "m/2"} and draws.count("m/10") > draws.count("m/2")
    p.add_canary_endpoint(dict(endpoint="fixed", weights=[1, 3], load_endpoints=["m/1", "m/2"]))
    assert p._canary_route["fixed"]["weights"] == [0.25, 0.75]
    assert p._process_canary("other") is None


def test_router_stall_during_reconfiguration():
    p = _router_with_fake()

    async def main():
        p._update_lock_flag = True

        async def release():
            await asyncio.sleep(0.05)
            p._update_lock_flag = False
        t = asyncio.ensure_future(release())
        r = await p.process_request("sum", None, [[2, 2]], "process")
        await t
        return r
    try:
        assert asyncio.run(main()).tolist() == [4.0]
    finally:
        p.shutdown()


def test_router_stats_sampling():
    p = _router_with_fake()
    seen = []
    p.set_stats_sink(seen.append, default_frequency=1.0)
    try:
        asyncio.run(p.process_request("sum", None, [[1, 1]], "process"))
    finally:
        p.shutdown()
    assert len(seen) == 1 and seen[0]["_url"] == "sum" and seen[0]["_count"] == 1 and "_latency" in seen[0]


def test_validate_partial_io_description_rejected():
    with pytest.raises(EndpointBackendEngineException, match="missing values"):
        ModelRequestProcessor._validate_model(
            ModelEndpoint(engine_type="b200", serving_url="m", input_type=["float32"]))


# ---------------------------------------------------------------- REST contract vs the reference app
class _IrisUser(object):
    def preprocess(self, body, state, collect_custom_statistics_fn=None):
        if "boom" in body:
            raise ValueError("bad request field")
        if "boom_rt" in body:
            raise RuntimeError("runtime failure")
        return [[body.get("x0"), body.get("x1"), body.get("x2"), body.get("x3")]]

    def postprocess(self, data, state, collect_custom_statistics_fn=None):
        return dict(y=data.tolist() if isinstance(data, np.ndarray) else data)


def _lr_fake_model(golden_dir):
    g = np.load(os.path.join(golden_dir, "lr_iris.npz"))
    W, b, classes = g["coef"], g["intercept"], g["classes"]
    return FakeModel(n_features=4, in_dtype=np.float64, out_dtype=np.int64,
                     fn=lambda x: classes[np.argmax(x @ W.T + b, axis=1)])


def test_rest_contract_matches_reference(golden_dir):
    from starlette.testclient import TestClient
    from clearml_serving_b200.main import create_app
    p = ModelRequestProcessor()
    for url, ver in (("iris", ""), ("iris/2", "2"), ("bad", "")):
        ep = ModelEndpoint(engine_type="b200", serving_url=url, version=ver)
        p._endpoints[url] = ep
        p._engine_processor_lookup[url] = make_fake_engine(ep, _lr_fake_model(golden_dir), preprocess=_IrisUser())
    client = TestClient(create_app(p), raise_server_exceptions=False)
    with open(os.path.join(golden_dir, "rest_contract.json")) as f:
        golden = json.load(f)
    try:
        for rec in golden:
            if rec.get("method") == "GET":
                resp = client.get(rec["path"])
            elif "gzip_json" in rec:
                resp = client.post(rec["path"], content=gzip.compress(json.dumps(rec["gzip_json"]).encode()),
                                   headers={"Content-Encoding": "gzip", "Content-Type": "application/json"})
            else:
                resp = client.post(rec["path"], json=rec["json"])
            assert resp.status_code == rec["status"], rec["name"]
            assert resp.json() == rec["response"], rec["name"]
    finally:
        p.shutdown()


# ---------------------------------------------------------------- replicas (multi-GPU router)
def test_round_robin_over_replicas():
    from clearml_serving_b200.router import parse_devices
    assert parse_devices({"b200.devices": "0,1,2,3"}, 0) == [0, 1, 2, 3]
    assert parse_devices({"b200.devices": [2, 5]}, 0) == [2, 5]
    assert parse_devices({"b200.device": 3}, 0) == [3]
    assert parse_devices(None, 1) == [1]
    ep = ModelEndpoint(engine_type="b200", serving_url="m")
    eng = make_fake_engine(ep, FakeModel(n_features=2), n_replicas=4)
    try:
        outs = [eng.process_sync([[i, i]]) for i in range(40)]
        assert [float(o[0]) for o in outs] == [2.0 * i for i in range(40)]
        st = eng.engine_stats()
        assert st["replicas"] == 4 and st["per_replica_requests"] == [10, 10, 10, 10]
    finally:
        eng.unload()


def test_model_hot_reload_drops_only_the_changed_endpoint(tmp_path):
    """sync_models(): the in-process counterpart of the Triton sidecar's model poll (triton_helper.py:91-194,226-289)
    + the engine drop of model_request_processor.py:1026-1028, per endpoint"""
    import joblib
    from sklearn.linear_model import LogisticRegression
    from clearml_serving_b200 import ModelEndpoint
    from clearml_serving_b200.model_repo import ModelRepository
    from clearml_serving_b200.model_request_processor import ModelRequestProcessor
    rng = np.random.default_rng(0)
    X = rng.standard_normal((80, 4))
    paths = []
    for k in range(2):
        p = tmp_path / "m{}.pkl".format(k)
        joblib.dump(LogisticRegression(max_iter=200).fit(X, (X[:, k] > 0).astype(int)), p)
        paths.append(str(p))
    proc = ModelRequestProcessor()
    for k, p in enumerate(paths):
        proc.add_endpoint(ModelEndpoint(engine_type="b200", serving_url="ep{}".format(k), model_id=p))

    class _Eng(object):
        unloaded = 0

        def unload(self):
            self.unloaded += 1
    engines = {"ep0": _Eng(), "ep1": _Eng()}
    proc._engine_processor_lookup.update(engines)
    repo = ModelRepository()
    assert proc.sync_models(repo) == []            # first pass: packs both, nothing to drop
    assert proc.sync_models(repo) == []            # unchanged files
    joblib.dump(LogisticRegression(max_iter=200).fit(X, (X[:, 2] > 0).astype(int)), paths[1])   # new model content
    assert proc.sync_models(repo) == ["ep1"]
    assert engines["ep1"].unloaded == 1 and engines["ep0"].unloaded == 0
    assert "ep1" not in proc._engine_processor_lookup and proc._engine_processor_lookup["ep0"] is engines["ep0"]
    # the daemon form: picks the change up by itself
    proc._engine_processor_lookup["ep1"] = engines["ep1"]
    proc.sync_models()                              # default repository: first pass
    proc.start_sync_daemon(poll_frequency_sec=0.05)
    joblib.dump(LogisticRegression(max_iter=200).fit(X, (X[:, 3] > 0).astype(int)), paths[1])
    deadline = time.time() + 5
    while "ep1" in proc._engine_processor_lookup and time.time() < deadline:
        time.sleep(0.02)
    proc.stop_sync_daemon()
    assert "ep1" not in proc._engine_processor_lookup and engines["ep1"].unloaded == 2
    proc._engine_processor_lookup.clear()


def test_canary_table_follows_endpoint_changes():
    """add_endpoint / remove_endpoint rebuild the canary route table (the reference rebuilds it on every
    configuration reload, model_request_processor.py:772-814)"""
    from clearml_serving_b200 import ModelEndpoint
    from clearml_serving_b200.model_request_processor import ModelRequestProcessor
    proc = ModelRequestProcessor()
    proc.add_canary_endpoint(dict(endpoint="pub", weights=[0.7, 0.3], load_endpoint_prefix="m/"))
    assert "pub" not in proc._canary_route                      # canary first, targets later
    proc.add_endpoint(ModelEndpoint(engine_type="custom", serving_url="m/1", version="1"))
    assert proc._canary_route["pub"]["endpoints"] == ["m/1"]
    proc.add_endpoint(ModelEndpoint(engine_type="custom", serving_url="m/2", version="2"))
    assert proc._canary_route["pub"]["endpoints"] == ["m/2", "m/1"] and proc._canary_route["pub"]["weights"] == [0.7, 0.3]
    proc.remove_endpoint("m/2")
    assert proc._canary_route["pub"]["endpoints"] == ["m/1"] and proc._canary_route["pub"]["weights"] == [1.0]
    proc.remove_endpoint("m/1")
    assert "pub" not in proc._canary_route
