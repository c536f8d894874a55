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


def xgb_legacy_binary_from_forest(forest, base_score=0.5, objective="reg:squarederror", binf=False, major=1, minor=7,
                                  attrs=True):
    """TEST-SIDE writer of XGBoost's legacy binary layout (xgboost 1.7 src/learner.cc, gbtree_model.h, tree_model.h;
    see the comment above formats.parse_xgboost_legacy_binary): what `Booster.save_model("xgb_model")` emits."""
    off = forest["tree_offset"]
    n_trees = len(off) - 1
    out = [b"binf"] if binf else []
    out.append(struct.pack("<fIiiiIII", base_score, int(forest["n_features"]), 0, 1 if attrs else 0, 0, major, minor, 1) + b"\0" * (136 - 32))
    for name in (objective, "gbtree"):
        out.append(struct.pack("<Q", len(name)) + name.encode())
    out.append(struct.pack("<iiiiqii", n_trees, 1, int(forest["n_features"]), 0, 0, 1, 0) + b"\0" * 128)
    for t in range(n_trees):
        s, e = int(off[t]), int(off[t + 1])
        n = e - s
        out.append(struct.pack("<iiiiii", 1, n, 0, 6, int(forest["n_features"]), 0) + b"\0" * 124)
        parent = np.full(n, -1, np.int32)
        for i in range(n):
            if forest["left"][s + i] >= 0:
                parent[forest["left"][s + i]] = i
                parent[forest["right"][s + i]] = i | (1 << 31) if False else i
        nodes = bytearray()
        for i in range(n):
            leaf = forest["left"][s + i] < 0
            sindex = 0 if leaf else (int(forest["feat"][s + i]) | (int(forest["default_left"][s + i]) << 31))
            info = forest["value"][s + i] if leaf else forest["thr"][s + i]
            nodes += struct.pack("<iiiIf", int(parent[i]), int(forest["left"][s + i]) if not leaf else -1,
                                 int(forest["right"][s + i]) if not leaf else 0, sindex, float(np.float32(info)))
        out.append(bytes(nodes))
        out.append(b"".join(struct.pack("<fffi", 0.0, 1.0, 0.0, 0) for _ in range(n)))
    out.append(struct.pack("<" + "i" * n_trees, *([0] * n_trees)))
    if attrs:   # vector<pair<string,string>>: the reader ignores everything after tree_info
        out.append(struct.pack("<Q", 1) + struct.pack("<Q", 14) + b"best_iteration" + struct.pack("<Q", 2) + b"99")
    return b"".join(out)


@pytest.mark.parametrize("binf", [False, True])
def test_xgboost_legacy_binary_model(tmp_path, binf):
    """examples/xgboost/train_model.py:28 saves `xgb_model` (no extension) = the legacy binary container."""
    from clearml_serving_b200 import model_repo
    f = orc.synth_xgb_forest(n_trees=13, depth=5, n_features=9, seed=4, ragged=True)
    raw = xgb_legacy_binary_from_forest(f, base_score=0.5, binf=binf)
    forest, base, link = formats.parse_xgboost_legacy_binary(raw)
    assert base == 0.5 and link == formats.LINK_IDENTITY and forest["n_features"] == 9
    rng = np.random.default_rng(3)
    X = rng.standard_normal((50, 9)).astype(np.float32)
    X[rng.random(X.shape) < 0.05] = np.nan
    want = orc.forest_predict_xgb(f, X, 0.5)
    assert np.array_equal(orc.forest_predict_xgb(forest, X, base), want)
    p = tmp_path / "xgb_model"
    p.write_bytes(raw)
    pm = model_repo.load_model(str(p))                      # sniffed, no extension, no framework tag
    assert np.array_equal(blob_interp.predict(pm.blob, X), want)
    assert model_repo.load_model(str(p), framework="XGBoost").blob == pm.blob
    with pytest.raises(ValueError, match="truncated"):
        formats.parse_xgboost_legacy_binary(raw[:len(raw) // 2])


def test_xgboost_logistic_objectives_and_base_score_forms():
    """binary:logitraw adds logit(base_score) (it used to add base_score itself); binary:logistic /
    reg:logistic additionally map the margin through xgboost's fp32 sigmoid; base_score may come as "[5E-1]"."""
    f = orc.synth_xgb_forest(n_trees=9, depth=3, n_features=5, seed=6)
    rng = np.random.default_rng(2)
    X = rng.standard_normal((30, 5)).astype(np.float32)
    for bs_text, bs in (("2.5E-1", 0.25), ("[7.5E-1]", 0.75), ("5E-1", 0.5)):
        margin0 = orc.xgb_prob_to_margin(bs)
        assert margin0 == pytest.approx(np.log(bs / (1 - bs)), abs=1e-6)
        doc = orc.xgb_json_from_forest(f, base_score=bs, objective="binary:logitraw")
        doc["learner"]["learner_model_param"]["base_score"] = bs_text
        forest, base, link = formats.parse_xgboost_json(doc)
        assert np.float32(base) == np.float32(margin0) and link == formats.LINK_IDENTITY
        want_margin = orc.forest_predict_xgb(f, X, margin0)
        assert np.array_equal(blob_interp.predict(formats.pack_xgboost_json(doc).blob, X), want_margin)
        for obj in ("binary:logistic", "reg:logistic"):
            doc["learner"]["objective"]["name"] = obj
            pm = formats.pack_xgboost_json(doc)
            assert pm.description["link"] == formats.LINK_SIGMOID
            got = blob_interp.predict(pm.blob, X)
            assert np.array_equal(got, orc.xgb_sigmoid(want_margin)) and np.all((got > 0) & (got < 1))
    raw = xgb_legacy_binary_from_forest(f, base_score=0.25, objective="binary:logistic")
    _forest, base, link = formats.parse_xgboost_legacy_binary(raw)
    assert np.float32(base) == np.float32(orc.xgb_prob_to_margin(0.25)) and link == formats.LINK_SIGMOID
    # a pre-1.0 writer stored the margin itself
    raw = xgb_legacy_binary_from_forest(f, base_score=-1.0986123, objective="binary:logistic", binf=True, major=0, minor=0)
    assert formats.parse_xgboost_legacy_binary(raw)[1] == pytest.approx(-1.0986123)


def test_xgboost_json_rejects_unsupported():
    f = orc.synth_xgb_forest(n_trees=2, depth=2, n_features=3, seed=0)
    doc = orc.xgb_json_from_forest(f, objective="multi:softprob")
    with pytest.raises(ValueError, match="objective"):
        formats.pack_xgboost_json(doc)
    doc = orc.xgb_json_from_forest(f, objective="binary:logistic", base_score=1.5)
    with pytest.raises(ValueError, match="base_score"):
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


# ---------------------------------------------------------------- LightGBM text models (reference engine "lightgbm", ps.py:486-501)
_LGB_MODEL = """tree
version=v3
num_class=1
num_tree_per_iteration=1
label_index=0
max_feature_idx=3
objective=regression
feature_names=a b c d
feature_infos=[-3:3] [-3:3] [-3:3] [-3:3]
tree_sizes=400 300 120

Tree=0
num_leaves=4
num_cat=0
split_feature=0 2 1
split_gain=10 5 3
threshold=0.5 -0.25 1.0000000180025095e-35
decision_type=2 10 8
left_child=1 -1 -3
right_child=2 -2 -4
leaf_value=0.125 -0.25 0.5 0.0625
leaf_weight=1 1 1 1
leaf_count=1 1 1 1
internal_value=0 0 0
internal_weight=0 0 0
internal_count=4 2 2
is_linear=0
shrinkage=1


Tree=1
num_leaves=3
num_cat=0
split_feature=3 0
split_gain=4 2
threshold=1.5 -1.75
decision_type=0 0
left_child=1 -1
right_child=-2 -3
leaf_value=0.03125 -0.015625 0.75
leaf_weight=1 1 1
leaf_count=1 1 1
internal_value=0 0
internal_weight=0 0
internal_count=3 2
is_linear=0
shrinkage=0.1


Tree=2
num_leaves=1
num_cat=0
leaf_value=0.001953125
is_linear=0
shrinkage=0.1


end of trees

feature_importances:
a=2
"""


def _lgb_reference_predict(text, X, average=False):
    """the published predictor, straight from the text: fval <= threshold in double, child >= 0 internal / < 0 leaf ~child,
    decision_type bit 1 default-left, bits 2-3 missing type (0: NaN reads as 0.0, 2: NaN takes the default side)"""
    trees = []
    for block in text.split("end of trees")[0].split("Tree=")[1:]:
        kv = dict(ln.split("=", 1) for ln in block.splitlines() if "=" in ln)
        trees.append({k: np.array(v.split(), dtype=np.float64) for k, v in kv.items()
                      if k in ("split_feature", "threshold", "decision_type", "left_child", "right_child", "leaf_value")})
    out = np.zeros(len(X), np.float64)
    for i, row in enumerate(X):
        acc = 0.0
        for t in trees:
            if "split_feature" not in t:
                acc += t["leaf_value"][0]
                continue
            node = 0
            while node >= 0:
                f, thr, dt = int(t["split_feature"][node]), t["threshold"][node], int(t["decision_type"][node])
                x = float(row[f])
                missing = (dt >> 2) & 3
                if np.isnan(x) and missing == 2:
                    left = bool(dt & 2)
                else:
                    left = (0.0 if np.isnan(x) else x) <= thr
                node = int(t["left_child"][node] if left else t["right_child"][node])
            acc += t["leaf_value"][~node]
        out[i] = acc / (len(trees) if average else 1)
    return out


def test_lightgbm_text_model_packs_to_the_published_semantics(tmp_path):
    from clearml_serving_b200 import formats, model_repo
    rng = np.random.default_rng(9)
    X = rng.standard_normal((200, 4)).astype(np.float32)
    X[rng.random(X.shape) < 0.15] = np.nan                     # both missing types are exercised
    X[:4] = [[0.5, 0, 0, 0], [np.float32(0.5000001), 1e-35, -0.25, 1.5], [0.5, 0, np.nextafter(np.float32(-0.25), np.float32(0)), 1.5],
             [np.nan, np.nan, np.nan, np.nan]]               # rows sitting exactly on thresholds
    p = tmp_path / "model.txt"
    p.write_text(_LGB_MODEL)
    pm = model_repo.load_model(str(p), framework="LightGBM")
    assert model_repo.load_model(str(p)).blob == pm.blob         # sniffed without the tag
    got = blob_interp.predict(pm.blob, X)
    want = _lgb_reference_predict(_LGB_MODEL, X)
    assert got.dtype == np.float64 and np.array_equal(got, want)
    # boosting=rf models average the trees
    rf = _LGB_MODEL.replace("objective=regression\n", "objective=regression\naverage_output\n")
    got_rf = blob_interp.predict(formats.pack_lightgbm_text(rf).blob, X)
    assert np.array_equal(got_rf, _lgb_reference_predict(rf, X, average=True))


@pytest.mark.parametrize("edit,msg", [
    (lambda t: t.replace("objective=regression", "objective=binary sigmoid:1"), "objective"),
    (lambda t: t.replace("num_class=1", "num_class=3"), "multi-class"),
    (lambda t: t.replace("decision_type=2 10 8", "decision_type=3 10 8"), "categorical"),
    (lambda t: t.replace("decision_type=2 10 8", "decision_type=6 10 8"), "zero_as_missing"),
    (lambda t: t.replace("split_feature=3 0", "split_feature=9 0"), "out of range"),
    (lambda t: t.replace("tree\nversion", "forest\nversion"), "not a LightGBM"),
])
def test_lightgbm_models_outside_the_supported_set_are_refused(edit, msg):
    from clearml_serving_b200 import formats
    with pytest.raises(ValueError, match=msg):
        formats.pack_lightgbm_text(edit(_LGB_MODEL))
