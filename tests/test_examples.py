"""examples/: the b200-ready counterparts of the reference's example endpoints.  CPU: the model
generators write files the ingestion layer packs, the user classes speak the reference's Preprocess surface, the REST app
serves them (host double underneath -- kernels are covered by the -m gpu suites)."""
import importlib.util
import io
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from clearml_serving_b200 import ModelEndpoint, formats, model_repo
from clearml_serving_b200.model_request_processor import ModelRequestProcessor
from tests import blob_interp
from tests.fakes import FakeModel, make_fake_engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX = os.path.join(ROOT, "examples")


def _user(folder):
    spec = importlib.util.spec_from_file_location("Preprocess_" + folder, os.path.join(EX, folder, "preprocess.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.Preprocess()


def _make(folder, out, *flags):
    r = subprocess.run([sys.executable, os.path.join(EX, folder, "make_model.py"), str(out)] + list(flags),
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]
    return r.stdout.strip().splitlines()[-1]


def test_sklearn_example_end_to_end_on_the_rest_app(tmp_path):
    path = _make("sklearn", tmp_path)
    pm = model_repo.load_model(path, framework="ScikitLearn")
    assert pm.kind is not None and pm.description["kind"] == "linear" and pm.description["n_features"] == 4
    from starlette.testclient import TestClient
    from clearml_serving_b200.main import create_app
    p = ModelRequestProcessor()
    ep = ModelEndpoint(engine_type="b200", serving_url="test_model_sklearn")
    p._endpoints["test_model_sklearn"] = ep
    p._engine_processor_lookup["test_model_sklearn"] = make_fake_engine(ep, FakeModel(n_features=4, in_dtype=np.float64),
                                                                        preprocess=_user("sklearn"))
    client = TestClient(create_app(p), raise_server_exceptions=False)
    try:
        r = client.post("/serve/test_model_sklearn", json={"x0": 1, "x1": 2, "x2": 3, "x3": 4})
        assert r.status_code == 200 and r.json() == {"y": [10.0]}          # the host double sums its row
    finally:
        p.shutdown()


def test_xgboost_example_model_and_missing_values(tmp_path):
    path = _make("xgboost", tmp_path, "--trees", "25")
    pm = model_repo.load_model(path, framework="XGBoost")
    user = _user("xgboost")
    row = user.preprocess({"x0": 0.25, "x5": -1.5, "x31": 2.0}, {}, None)
    assert row.dtype == np.float32 and row.shape == (1, 32) and np.isnan(row[0, 1]) and row[0, 31] == 2.0
    y = blob_interp.predict(pm.blob, row)                                   # missing features follow default_left
    with open(path) as f:
        trees = json.load(f)["learner"]["gradient_booster"]["model"]["trees"]
    acc = np.float32(0.5)
    for t in trees:
        n = 0
        while t["left_children"][n] != -1:
            x = row[0, t["split_indices"][n]]
            left = bool(t["default_left"][n]) if np.isnan(x) else bool(x < np.float32(t["split_conditions"][n]))
            n = t["left_children"][n] if left else t["right_children"][n]
        acc = np.float32(acc + np.float32(t["split_conditions"][n]))
    assert y.shape == (1,) and y[0] == acc
    assert user.postprocess(y, {}, None) == {"y": [float(acc)]}


def test_pytorch_example_decodes_images_and_packs(tmp_path):
    from PIL import Image
    user = _user("pytorch")
    img = Image.fromarray(np.random.default_rng(0).integers(0, 256, (60, 80, 3), dtype=np.uint8).astype(np.uint8))
    buf = io.BytesIO()
    img.save(buf, format="PNG")
    x = user.preprocess(buf.getvalue(), {}, None)
    assert x.dtype == np.uint8 and x.shape == (1, 3, 224, 224) and x.flags.c_contiguous
    with pytest.raises(RuntimeError):
        user.preprocess(b"not an image", {}, None)                           # -> 500 like the reference's example
    two = user.preprocess({"pixels": np.zeros((2, 3, 224, 224), np.uint8).tolist()}, {}, None)
    assert two.shape == (2, 3, 224, 224)
    assert user.postprocess(np.array([[0.1, 0.7, 0.2], [0.9, 0.0, 0.1]]), {}, None) == {"class": [1, 0]}
    path = _make("pytorch", tmp_path, "--arch", "resnet18")
    pm = model_repo.load_model(path, framework="PyTorch")
    assert pm.description["arch"] == "resnet" and pm.description["num_classes"] == 1000


def test_huggingface_example_ids_and_model_folder(tmp_path):
    user = _user("huggingface")
    ids, types, mask = user.preprocess({"input_ids": list(range(5, 25))}, {}, None)
    assert ids == [list(range(5, 25))] and types == [[0] * 20] and mask == [[1] * 20]
    assert len(user.preprocess({"input_ids": [1] * 1000}, {}, None)[0][0]) == 256   # truncated like max_length
    with pytest.raises(ValueError):
        user.preprocess({"text": "no tokenizer folder configured"}, {}, None)
    folder = _make("huggingface", tmp_path, "--tiny")
    pm = model_repo.load_model(folder)
    assert pm.description["kind"] == "graph" and pm.description["num_labels"] == 2


def test_endpoints_file_is_a_valid_endpoint_table():
    with open(os.path.join(EX, "endpoints.json")) as f:
        cfg = json.load(f)
    urls = set()
    for d in cfg["endpoints"]:
        ep = ModelEndpoint(**d)
        assert ep.engine_type == "b200" and os.path.exists(os.path.join(ROOT, ep.preprocess_artifact))
        ModelRequestProcessor._validate_model(ep)
        urls.add(ep.serving_url)
    assert urls == {"test_model_sklearn", "test_model_xgb", "test_model_pytorch", "transformer_model"}   # the reference's endpoint names
