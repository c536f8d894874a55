"""formats.py (model ingestion) on CPU: the packed blob, executed by tests/blob_interp.py with the
kernel's rules, must reproduce the oracle / the reference goldens bit for bit."""
import json
import os
import struct

import numpy as np
import pytest

from clearml_serving_b200 import formats
from oracle import oracle as orc
from tests import blob_interp


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name)))


@pytest.mark.parametrize("name", ["sk_gbr.npz", "sk_rf.npz"])
def test_sklearn_blob_matches_reference_golden(golden_dir, name):
    g = _load(golden_dir, name)
    pm = formats.pack_forest(g, "skl", base=float(g["init"]), scale=float(g["scale"]), divisor=float(g["divisor"]))
    y = blob_interp.predict(pm.blob, g["X"][:160])  # includes the rows sitting on thresholds
    assert np.array_equal(y, g["y"][:160])


@pytest.mark.parametrize("ragged", [False, True])
def test_xgb_blob_matches_oracle(ragged):
    f = orc.synth_xgb_forest(n_trees=23, depth=6, n_features=32, seed=11, ragged=ragged)
    rng = np.random.default_rng(1)
    X = rng.standard_normal((64, 32)).astype(np.float32)
    X[rng.random(X.shape) < 0.05] = np.nan
    pm = formats.pack_forest(f, "xgb", base=0.5)
    assert np.array_equal(blob_interp.predict(pm.blob, X), orc.forest_predict_xgb(f, X, 0.5))


def test_xgboost_json_roundtrip(tmp_path):
    f = orc.synth_xgb_forest(n_trees=11, depth=4, n_features=7, seed=2, ragged=True)
    doc = orc.xgb_json_from_forest(f, base_score=0.25)
    p = tmp_path / "model.json"
    p.write_text(json.dumps(doc))
    rng = np.random.default_rng(0)
    X = rng.standard_normal((40, 7)).astype(np.float32)
    want = orc.forest_predict_xgb(f, X, 0.25)
    for src in (str(p), doc, json.dumps(doc)):
        pm = formats.pack_xgboost_json(src)
        assert np.array_equal(blob_interp.predict(pm.blob, X), want)
    pm = formats.load_model_file(str(p))
    assert pm.description["n_trees"] == 11 and pm.description["mode"] == "xgb"


def test_xgboost_json_rejects_unsupported():
    f = orc.synth_xgb_forest(n_trees=2, depth=2, n_features=3, seed=0)
    doc = orc.xgb_json_from_forest(f, objective="binary:logistic")
    with pytest.raises(ValueError, match="objective"):
        formats.pack_xgboost_json(doc)
    doc = orc.xgb_json_from_forest(f)
    doc["learner"]["learner_model_param"]["num_class"] = "3"
    with pytest.raises(ValueError, match="multi-class"):
        formats.pack_xgboost_json(doc)


def test_threshold_conversion_is_exact():
    rng = np.random.default_rng(3)
    thr = np.concatenate([rng.standard_normal(2000) * 10, [0.0, -0.0, 1.0, 0.1, 3.4e38, -3.4e38, 1e-45, 5e-324]])
    t32 = formats._skl_threshold_to_f32_strict(thr)
    for t64, t in zip(thr, t32):
        below = np.nextafter(t, np.float32(-np.inf))
        assert np.float64(below) <= t64 < np.float64(t) or (np.isinf(t) and np.float64(below) <= t64)


def test_sklearn_estimators_pack(golden_dir):
    from sklearn.ensemble import GradientBoostingRegressor, RandomForestRegressor
    from sklearn.linear_model import LogisticRegression
    from sklearn.tree import DecisionTreeRegressor
    rng = np.random.default_rng(0)
    X = rng.standard_normal((200, 6))
    y = X[:, 0] - 2 * X[:, 1] * X[:, 2]
    Xq = rng.standard_normal((64, 6)).astype(np.float32)
    for est in (GradientBoostingRegressor(n_estimators=12, max_depth=3, random_state=0),
                RandomForestRegressor(n_estimators=7, max_depth=5, random_state=0, n_jobs=1),
                DecisionTreeRegressor(max_depth=6, random_state=0)):
        est.fit(X, y)
        pm = formats.pack_sklearn(est)
        got = blob_interp.predict(pm.blob, Xq)
        assert np.array_equal(got, est.predict(Xq)), type(est).__name__
    lr = LogisticRegression(max_iter=500).fit(X, (y > 0).astype(int))
    pm = formats.pack_sklearn(lr)
    assert pm.kind == 2 and pm.description["n_out"] == 1
    magic, ver, nf, no, nc, _ = struct.unpack_from("<4sIIIII", pm.blob, 0)
    assert (magic, ver, nf, no, nc) == (b"B2SL", 1, 6, 1, 2)


def test_pack_forest_rejects_bad_trees():
    f = orc.synth_xgb_forest(n_trees=2, depth=2, n_features=3, seed=0)
    bad = dict(f)
    bad["feat"] = f["feat"].copy()
    bad["feat"][0] = 9
    with pytest.raises(ValueError, match="feature index"):
        formats.pack_forest(bad, "xgb")
    bad = dict(f)
    bad["left"] = f["left"].copy()
    bad["left"][0] = 100
    with pytest.raises(ValueError, match="out of range"):
        formats.pack_forest(bad, "xgb")
    with pytest.raises(ValueError):
        formats.pack_forest(f, "lightgbm")


def test_stem_space_to_depth_weight_is_the_same_convolution():
    """formats.stem_s2d_weight + the z layout of csrc/conv.cu nchw_to_s2d_kernel, executed in numpy fp64, reproduce
    torch's 7x7 stride-2 pad-3 convolution (the ResNet stem the engine runs as a 4x4 stride-1 implicit GEMM)"""
    import torch
    from clearml_serving_b200 import formats
    rng = np.random.default_rng(5)
    for (n, C, H, W, Cout) in [(2, 3, 20, 24, 8), (1, 3, 15, 13, 16), (1, 1, 8, 8, 8)]:
        x = rng.standard_normal((n, C, H, W))
        w = rng.standard_normal((Cout, C, 7, 7))
        ref = torch.nn.functional.conv2d(torch.from_numpy(x), torch.from_numpy(w), stride=2, padding=3).numpy()
        OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        Hz, Wz = OH + 3, OW + 3
        xp = np.zeros((n, 4, 2 * Hz, 2 * Wz))
        xp[:, :C, 3:3 + H, 3:3 + W] = x
        # z[n, hz, wz, (dy, dx, c)]
        z = xp.reshape(n, 4, Hz, 2, Wz, 2).transpose(0, 2, 4, 3, 5, 1).reshape(n, Hz, Wz, 16)
        w2 = formats.stem_s2d_weight(w)
        assert w2.shape == (Cout, 256)
        got = np.zeros((n, OH, OW, Cout))
        for a in range(4):
            for b in range(4):
                got += z[:, a:a + OH, b:b + OW, :] @ w2[:, a * 64 + b * 16:a * 64 + b * 16 + 16].T
        np.testing.assert_allclose(got.transpose(0, 3, 1, 2), ref, rtol=1e-10, atol=1e-10)
