"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the oracle on the same
seeded inputs and against the golden fixtures recorded from the reference.  Bit-exact everywhere
(tree sums are sequential fp32 / fp64 by construction; labels are integers)."""
import os

import numpy as np
import pytest

from clearml_serving_b200 import formats
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name)))


def _run_batch(native, model, stream, reqs):
    ev, outs, keep = stream.infer_batch([[r] for r in reqs])
    stream.wait(ev)
    return [o[0] for o in outs]


@pytest.fixture(scope="module")
def cfg2(gpu_native):
    """BASELINE.json configs[1]: 1000 trees x depth 6 x 32 features, base_score 0.5."""
    forest = orc.synth_xgb_forest(n_trees=1000, depth=6, n_features=32, seed=0)
    pm = formats.pack_forest(forest, "xgb", base=0.5)
    model = gpu_native.Model(pm.kind, pm.blob, device=0)
    yield forest, model
    model.free()


def test_cfg2_batch64_bit_exact(gpu_native, cfg2):
    forest, model = cfg2
    assert int(model.info.algo_bytes_fixed) == 760000 and int(model.info.algo_bytes_per_row) == 132
    rng = np.random.default_rng(1)
    X = rng.standard_normal((64, 32)).astype(np.float32)
    want = orc.forest_predict_xgb(forest, X, 0.5)
    st = gpu_native.Stream(model, 64, 0, 4)
    try:
        got = np.concatenate(_run_batch(gpu_native, model, st, [X[i:i + 1] for i in range(64)]))
        assert got.dtype == np.float32 and np.array_equal(got, want)
        # same rows, different batch compositions => same bits (no dependence on batch-mates)
        for split in ([64], [1, 63], [7, 25, 32], [13] * 4 + [12]):
            parts, o = [], 0
            for n in split:
                parts.append(X[o:o + n]); o += n
            got2 = np.concatenate(_run_batch(gpu_native, model, st, parts))
            assert np.array_equal(got2, want)
    finally:
        st.destroy()


@pytest.mark.parametrize("n_rows", [1, 2, 31, 32, 33, 65, 257, 1000])
def test_cfg2_ragged_batch_sizes_with_nan(gpu_native, cfg2, n_rows):
    forest, model = cfg2
    rng = np.random.default_rng(100 + n_rows)
    X = rng.standard_normal((n_rows, 32)).astype(np.float32)
    X[rng.random(X.shape) < 0.01] = np.nan   # 1 % missing values -> default directions
    want = orc.forest_predict_xgb(forest, X, 0.5)
    st = gpu_native.Stream(model, max(n_rows, 1), 0, 2)
    try:
        got = _run_batch(gpu_native, model, st, [X])[0]
        assert np.array_equal(got, want)
    finally:
        st.destroy()


@pytest.mark.parametrize("env", [{"B2S_FOREST_WIDE": "0"}, {"B2S_FOREST_WIDE": "0", "B2S_FOREST_NO_STAGED": "1"},
                                 {"B2S_FOREST_WIDE_C": "8"}, {"B2S_FOREST_WIDE_C": "4"}])
def test_cfg2_other_kernel_forms_same_bits(gpu_native, monkeypatch, env):
    """the fallback kernels (staged slices / plain cluster: forests the compact images cannot hold) and the wide
    kernel at smaller cluster sizes (2 / 4 rows per owner rank) produce the same bits as the default form"""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    forest = orc.synth_xgb_forest(n_trees=1000, depth=6, n_features=32, seed=0)
    pm = formats.pack_forest(forest, "xgb", base=0.5)
    model = gpu_native.Model(pm.kind, pm.blob, device=0)     # the switches are read at model load
    rng = np.random.default_rng(21)
    st = gpu_native.Stream(model, 1500, 0, 2)
    try:
        for n in (1, 17, 64, 300, 1500):
            X = rng.standard_normal((n, 32)).astype(np.float32)
            X[rng.random(X.shape) < 0.01] = np.nan
            assert np.array_equal(_run_batch(gpu_native, model, st, [X])[0], orc.forest_predict_xgb(forest, X, 0.5))
    finally:
        st.destroy(); model.free()


def test_cfg2_large_batch_rows_kernel_and_device_path(gpu_native, cfg2):
    """> 4096 rows takes the one-row-per-thread kernel; also exercises b2s_infer_device."""
    forest, model = cfg2
    n = 20000
    rng = np.random.default_rng(7)
    X = rng.standard_normal((n, 32)).astype(np.float32)
    X[rng.random(X.shape) < 0.01] = np.nan
    want = orc.forest_predict_xgb(forest, X, 0.5, n_threads=0)
    st = gpu_native.Stream(model, n, 0, 1)
    d_in = gpu_native.DeviceBuffer(X.nbytes)
    d_out = gpu_native.DeviceBuffer(n * 4)
    try:
        d_in.upload(X)
        st.infer_device(n, [d_in.ptr], [d_out.ptr])
        st.synchronize()
        assert np.array_equal(d_out.download(np.float32, n), want)
        # and the cluster kernel at its upper edge (128 clusters), through the device path too
        st.infer_device(4096, [d_in.ptr], [d_out.ptr])
        st.synchronize()
        assert np.array_equal(d_out.download(np.float32, 4096), want[:4096])
        # linearity-free size-independent property: permuting rows permutes outputs
        perm = rng.permutation(n)
        d_in.upload(X[perm])
        st.infer_device(n, [d_in.ptr], [d_out.ptr])
        st.synchronize()
        assert np.array_equal(d_out.download(np.float32, n), want[perm])
    finally:
        d_in.free(); d_out.free(); st.destroy()


@pytest.mark.parametrize("shape", [dict(n_trees=1, depth=1, n_features=1), dict(n_trees=3, depth=10, n_features=5),
                                    dict(n_trees=257, depth=4, n_features=100), dict(n_trees=40, depth=3, n_features=2000),
                                    dict(n_trees=2500, depth=3, n_features=7)])
def test_forest_shapes_edge_cases(gpu_native, shape):
    """single stump, deep ragged trees, tree count not a multiple of the cluster tile, a feature
    count too wide for the shared-memory x tile (global-memory fallback path), and a forest larger
    than one leaf-matrix chunk (multi-chunk loop with the accumulator carried across chunks)."""
    forest = orc.synth_xgb_forest(seed=5, ragged=shape["depth"] > 4, **shape)
    pm = formats.pack_forest(forest, "xgb", base=-1.25)
    model = gpu_native.Model(pm.kind, pm.blob, device=0)
    rng = np.random.default_rng(3)
    F = shape["n_features"]
    st = gpu_native.Stream(model, 100, 0, 2)
    try:
        for n in (1, 45, 100):
            X = rng.standard_normal((n, F)).astype(np.float32)
            X[rng.random(X.shape) < 0.05] = np.nan
            got = _run_batch(gpu_native, model, st, [X])[0]
            assert np.array_equal(got, orc.forest_predict_xgb(forest, X, -1.25))
    finally:
        st.destroy(); model.free()


@pytest.mark.parametrize("name", ["sk_gbr.npz", "sk_rf.npz"])
def test_sklearn_goldens_bit_exact(gpu_native, golden_dir, name):
    """fp64 mode against outputs of the reference's SKLearnPreprocessRequest.process."""
    g = _load(golden_dir, name)
    pm = formats.pack_forest(g, "skl", base=float(g["init"]), scale=float(g["scale"]), divisor=float(g["divisor"]))
    model = gpu_native.Model(pm.kind, pm.blob, device=0)
    st = gpu_native.Stream(model, 512, 0, 2)
    try:
        X = g["X"]
        got = np.concatenate(_run_batch(gpu_native, model, st, [X[:100], X[100:101], X[101:]]))
        assert got.dtype == np.float64 and np.array_equal(got, g["y"])
    finally:
        st.destroy(); model.free()


def test_cfg2_shape_golden_bit_exact(gpu_native, golden_dir):
    """BASELINE.json configs[1] SHAPE (1000 trees x depth 6 x 32 features), outputs of the reference's own
    SKLearnPreprocessRequest.process: the GPU fp64 forest must reproduce every bit, as one batch, as single-row
    requests collated by the C ABI (max_batch 64, the serving shape) and on rows sitting on thresholds +-1 ulp."""
    g = _load(golden_dir, "sk_gbr_cfg2.npz")
    pm = formats.pack_forest(g, "skl", base=float(g["init"]), scale=float(g["scale"]), divisor=float(g["divisor"]))
    model = gpu_native.Model(pm.kind, pm.blob, device=0)
    st = gpu_native.Stream(model, 320, 0, 2)
    try:
        X = g["X"]
        got = _run_batch(gpu_native, model, st, [X])[0]
        assert got.dtype == np.float64 and np.array_equal(got, g["y"])
        for s in range(0, 320, 64):
            got = np.concatenate(_run_batch(gpu_native, model, st, [X[i:i + 1] for i in range(s, s + 64)]))
            assert np.array_equal(got, g["y"][s:s + 64])
    finally:
        st.destroy(); model.free()


def test_sklearn_large_forest_fp64_rows_kernel(gpu_native):
    from sklearn.ensemble import GradientBoostingRegressor
    rng = np.random.default_rng(0)
    Xtr = rng.standard_normal((400, 10))
    ytr = Xtr[:, 0] * Xtr[:, 1] + np.sin(Xtr[:, 2])
    # 700 stages: more than one fp64 leaf-matrix chunk (512 trees) in the cluster kernel
    est = GradientBoostingRegressor(n_estimators=700, max_depth=2, random_state=0).fit(Xtr, ytr)
    pm = formats.pack_sklearn(est)
    model = gpu_native.Model(pm.kind, pm.blob, device=0)
    n = 9000
    X = rng.standard_normal((n, 10)).astype(np.float32)
    st = gpu_native.Stream(model, n, 0, 1)
    try:
        got = _run_batch(gpu_native, model, st, [X])[0]
        assert np.array_equal(got, est.predict(X))   # sklearn itself, same box
        got_small = _run_batch(gpu_native, model, st, [X[:77]])[0]
        assert np.array_equal(got_small, got[:77])
    finally:
        st.destroy(); model.free()


def test_linear_golden_labels(gpu_native, golden_dir):
    g = _load(golden_dir, "lr_iris.npz")
    for suffix in ("", "_b"):
        pm = formats.pack_linear(g["coef" + suffix], g["intercept" + suffix], g["classes" + suffix])
        model = gpu_native.Model(pm.kind, pm.blob, device=0)
        st = gpu_native.Stream(model, 256, 0, 2)
        try:
            ev, outs, keep = st.infer_batch([[g["X"][:128]], [g["X"][128:]]])
            st.wait(ev)
            labels = np.concatenate([o[0] for o in outs])
            scores = np.concatenate([o[1] for o in outs])
            assert np.array_equal(labels, g["y" + suffix])
            o_scores, _ = orc.linear_predict(g["X"], g["coef" + suffix], g["intercept" + suffix])
            assert np.array_equal(scores.reshape(o_scores.shape), o_scores)   # un-fused fp64, same order
        finally:
            st.destroy(); model.free()


def test_staged_slot_api_and_pipelining(gpu_native, cfg2):
    forest, model = cfg2
    rng = np.random.default_rng(11)
    st = gpu_native.Stream(model, 64, 0, 3)
    try:
        pending = []
        for k in range(3):
            X = rng.standard_normal((10 + k, 32)).astype(np.float32)
            slot = st.acquire()
            slot.inputs[0][:X.shape[0]] = X
            pending.append((st.submit(slot, X.shape[0]), slot, X))
        with pytest.raises(gpu_native.B2SError) as ei:
            st.acquire()                                      # all 3 slots in flight
        assert ei.value.code == gpu_native.B2S_ERR_BUSY
        for ev, slot, X in pending:
            st.wait(ev)
            assert np.array_equal(slot.outputs[0][:X.shape[0]], orc.forest_predict_xgb(forest, X, 0.5))
            st.release(slot)
        assert st.query(pending[0][0]) is True
    finally:
        st.destroy()


def test_abi_error_paths(gpu_native, cfg2):
    forest, model = cfg2
    st = gpu_native.Stream(model, 8, 0, 2)
    try:
        with pytest.raises(gpu_native.B2SError, match="multiple of the model"):
            st.infer_batch([[np.zeros((1, 31), np.float32)]])
        with pytest.raises(gpu_native.B2SError, match="dtype"):
            st.infer_batch([[np.zeros((1, 32), np.float64)]])
        with pytest.raises(gpu_native.B2SError, match="exceed"):
            st.infer_batch([[np.zeros((9, 32), np.float32)]])
        with pytest.raises(gpu_native.B2SError, match="bad magic"):
            gpu_native.Model(gpu_native.MODEL_FOREST, b"\0" * 256, device=0)
        with pytest.raises(gpu_native.B2SError, match="still references"):
            gpu_native.check(gpu_native.lib().b2s_model_free(model.handle))
        # the failures above must not leak slots
        X = np.ones((8, 32), np.float32)
        for _ in range(5):
            assert np.array_equal(_run_batch(gpu_native, model, st, [X])[0], orc.forest_predict_xgb(forest, X, 0.5))
    finally:
        st.destroy()


def test_two_endpoints_two_streams_concurrently(gpu_native, cfg2, golden_dir):
    """multi-model per GPU: one CUDA stream per endpoint over the shared pinned arena."""
    forest, model = cfg2
    g = _load(golden_dir, "sk_gbr.npz")
    pm = formats.pack_forest(g, "skl", base=float(g["init"]), scale=float(g["scale"]), divisor=1.0)
    model2 = gpu_native.Model(pm.kind, pm.blob, device=0)
    s1 = gpu_native.Stream(model, 64, 0, 4)
    s2 = gpu_native.Stream(model2, 64, 0, 4)
    rng = np.random.default_rng(2)
    try:
        X1 = rng.standard_normal((4, 64, 32)).astype(np.float32)
        evs = []
        for k in range(4):
            evs.append((s1, s1.infer_batch([[X1[k]]])))
            evs.append((s2, s2.infer_batch([[g["X"][k * 64:(k + 1) * 64]]])))
        for idx, (s, (ev, outs, keep)) in enumerate(evs):
            s.wait(ev)
            k = idx // 2
            if s is s1:
                assert np.array_equal(outs[0][0], orc.forest_predict_xgb(forest, X1[k], 0.5))
            else:
                assert np.array_equal(outs[0][0], g["y"][k * 64:(k + 1) * 64])
    finally:
        s1.destroy(); s2.destroy(); model2.free()


def test_lightgbm_text_model_bit_exact(gpu_native):
    """LightGBM text model (the reference's `lightgbm` engine, preprocess_service.py:486-501) on the fp64 forest kernel:
    bit-identical to the restatement of the published predictor (tests/test_formats.py), NaN rows of both missing types
    and rows sitting on thresholds included, plain and `average_output`"""
    from tests.test_formats import _LGB_MODEL, _lgb_reference_predict
    rng = np.random.default_rng(9)
    X = rng.standard_normal((300, 4)).astype(np.float32)
    X[rng.random(X.shape) < 0.15] = np.nan
    X[:3] = [[0.5, 0, 0, 0], [np.float32(0.5000001), 1e-35, -0.25, 1.5], [np.nan, np.nan, np.nan, np.nan]]
    for text, avg in ((_LGB_MODEL, False), (_LGB_MODEL.replace("objective=regression\n", "objective=regression\naverage_output\n"), True)):
        pm = formats.pack_lightgbm_text(text)
        model = gpu_native.Model(pm.kind, pm.blob, device=0)
        stream = gpu_native.Stream(model, 512, 0, 2)
        try:
            got = np.concatenate(_run_batch(gpu_native, model, stream, [X[:1], X[1:65], X[65:]]))
            want = _lgb_reference_predict(text, X, average=avg)
            assert got.dtype == np.float64 and np.array_equal(got, want)
        finally:
            stream.destroy()
            model.free()
