"""GPU engine-level tests (-m gpu): the real B200PreprocessRequest (plugin API) and the REST app on
top of libb200serve, against the goldens recorded from the reference."""
import asyncio
import gzip
import json
import os

import numpy as np
import pytest

from clearml_serving_b200 import BasePreprocessRequest, ModelEndpoint, formats
from clearml_serving_b200.model_request_processor import ModelRequestProcessor
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


class _IrisUser(object):
    def preprocess(self, body, state, collect_custom_statistics_fn=None):
        if "boom" in body:
            raise ValueError("bad request field")
        if "boom_rt" in body:
            raise RuntimeError("runtime failure")
        return [[body.get("x0"), body.get("x1"), body.get("x2"), body.get("x3")]]

    def postprocess(self, data, state, collect_custom_statistics_fn=None):
        return dict(y=data.tolist() if isinstance(data, np.ndarray) else data)


def _engine(ep, preprocess=None):
    cls = BasePreprocessRequest.get_engine_cls("b200")
    e = cls(model_endpoint=ep, task=None)
    e._preprocess = preprocess
    return e


def test_xgboost_json_endpoint_through_plugin_api(gpu_native, tmp_path):
    """configs[1] end to end: XGBoost-schema JSON file -> endpoint -> concurrent async requests."""
    forest = orc.synth_xgb_forest(n_trees=1000, depth=6, n_features=32, seed=0)
    path = tmp_path / "xgb_model.json"
    path.write_text(json.dumps(orc.xgb_json_from_forest(forest, base_score=0.5)))
    ep = ModelEndpoint(engine_type="b200", serving_url="test_model_xgb", model_id=str(path),
                       auxiliary_cfg={"max_batch_size": 64, "dynamic_batching.max_queue_delay_microseconds": 1000})
    eng = _engine(ep)
    rng = np.random.default_rng(1)
    X = rng.standard_normal((500, 32)).astype(np.float32)
    X[rng.random(X.shape) < 0.01] = np.nan
    want = orc.forest_predict_xgb(forest, X, 0.5)

    async def main():
        async def one(i):
            body = [[None if np.isnan(v) else float(v) for v in X[i]]]     # what a JSON body decodes to
            return await eng.process(body, {}, None)
        return await asyncio.gather(*[one(i) for i in range(len(X))])

    try:
        outs = asyncio.run(main())
        got = np.concatenate(outs)
        assert got.dtype == np.float32 and np.array_equal(got, want)
        st = eng.engine_stats()
        assert st["requests"] == 500 and st["max_batch_rows"] <= 64 and st["batches"] < 500
        # blocking variant, multi-row request
        assert np.array_equal(eng.process_sync(X[:17]), want[:17])
        with pytest.raises(ValueError, match="exceeds max_batch_size"):
            eng.process_sync(X[:65])
    finally:
        eng.unload()


def test_sklearn_pickle_endpoint(gpu_native, tmp_path):
    import joblib
    from sklearn.ensemble import RandomForestRegressor
    rng = np.random.default_rng(0)
    Xtr = rng.standard_normal((300, 8))
    est = RandomForestRegressor(n_estimators=20, max_depth=6, random_state=0, n_jobs=1).fit(Xtr, Xtr[:, 0] * Xtr[:, 1])
    path = tmp_path / "sklearn-model.pkl"
    joblib.dump(est, path)
    eng = _engine(ModelEndpoint(engine_type="b200", serving_url="rf", model_id=str(path)))
    try:
        Xq = rng.standard_normal((40, 8)).astype(np.float32)
        got = eng.process_sync(Xq.tolist())
        assert got.dtype == np.float64 and np.array_equal(got, est.predict(Xq))
    finally:
        eng.unload()


def test_rest_contract_with_gpu_engine(gpu_native, golden_dir):
    """configs[0] through REST: identical status codes, bodies and detail strings as the reference
    FastAPI app (tests/golden/rest_contract.json), with the linear model running on the GPU."""
    from starlette.testclient import TestClient
    from clearml_serving_b200.main import create_app
    g = np.load(os.path.join(golden_dir, "lr_iris.npz"))
    packed = formats.pack_linear(g["coef"], g["intercept"], g["classes"])

    class _User(_IrisUser):
        def load(self, local_file_name):
            return packed

    p = ModelRequestProcessor()
    cls = BasePreprocessRequest.get_engine_cls("b200")
    for url, ver in (("iris", ""), ("iris/2", "2"), ("bad", "")):
        ep = ModelEndpoint(engine_type="b200", serving_url=url, version=ver)
        p._endpoints[url] = ep
        e = cls.__new__(cls)
        BasePreprocessRequest.__init__(e, model_endpoint=ep, task=None)
        e._preprocess = _User()
        e._model = packed
        e._b200_setup()
        p._engine_processor_lookup[url] = e
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
        # the full golden request stream, through the router
        async def stream():
            out = []
            for row in g["X"]:
                body = {"x%d" % i: float(v) for i, v in enumerate(row)}
                r = await p.process_request("iris", None, body, "process")
                out.append(r["y"][0])
            return out
        assert np.array_equal(np.asarray(asyncio.run(stream())), g["y"])
    finally:
        p.shutdown()
