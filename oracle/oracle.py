"""ctypes front-end of the CPU oracle (oracle/forest_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

A forest in "source form" is a dict of numpy arrays (the common denominator of the XGBoost JSON
schema and sklearn's tree_ arrays):
    tree_offset int32[n_trees+1], left/right/feat int32[n_nodes] (children LOCAL to their tree,
    left<0 => leaf), thr float64[n_nodes], default_left uint8[n_nodes], value float64[n_nodes],
    n_features int
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class _Forest(ctypes.Structure):
    _fields_ = [
        ("n_trees", ctypes.c_int32),
        ("n_features", ctypes.c_int32),
        ("tree_offset", ctypes.c_void_p),
        ("left", ctypes.c_void_p),
        ("right", ctypes.c_void_p),
        ("feat", ctypes.c_void_p),
        ("thr", ctypes.c_void_p),
        ("default_left", ctypes.c_void_p),
        ("value", ctypes.c_void_p),
    ]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "forest_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.oracle_max_threads.restype = ctypes.c_int
    return _LIB


def max_threads():
    return int(lib().oracle_max_threads())


def _keep(forest):
    arrs = dict(
        tree_offset=np.ascontiguousarray(forest["tree_offset"], dtype=np.int32),
        left=np.ascontiguousarray(forest["left"], dtype=np.int32),
        right=np.ascontiguousarray(forest["right"], dtype=np.int32),
        feat=np.ascontiguousarray(forest["feat"], dtype=np.int32),
        thr=np.ascontiguousarray(forest["thr"], dtype=np.float64),
        default_left=np.ascontiguousarray(forest["default_left"], dtype=np.uint8),
        value=np.ascontiguousarray(forest["value"], dtype=np.float64),
    )
    f = _Forest()
    f.n_trees = int(len(arrs["tree_offset"]) - 1)
    f.n_features = int(forest["n_features"])
    for k, a in arrs.items():
        setattr(f, k, a.ctypes.data)
    return f, arrs


class ForestHandle(object):
    """Pre-compiled forest for timing loops: compact AoS nodes (oracle_forest_compile), no per-call
    array conversion.  Bit-identical to forest_predict_xgb (tests/test_oracle.py)."""

    def __init__(self, forest):
        f, keep = _keep(forest)
        L = lib()
        L.oracle_forest_compile.restype = ctypes.c_void_p
        L.oracle_forest_compile.argtypes = [ctypes.c_void_p]
        L.oracle_compiled_free.argtypes = [ctypes.c_void_p]
        L.oracle_compiled_free.restype = None
        self._fn = L.oracle_compiled_predict_xgb
        self._fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_float,
                             ctypes.c_void_p, ctypes.c_int]
        self._fn.restype = None
        self.n_features = int(forest["n_features"])
        self._c = L.oracle_forest_compile(ctypes.addressof(f))
        del keep

    def predict_xgb_into(self, X, base_score, out, n_threads=1):
        """X: C-contiguous float32 [n, F]; out: float32 [n]."""
        self._fn(self._c, X.ctypes.data, X.shape[0], base_score, out.ctypes.data, n_threads)
        return out

    def __del__(self):
        try:
            if getattr(self, "_c", None):
                lib().oracle_compiled_free(self._c)
                self._c = None
        except Exception:  # noqa
            pass


def forest_predict_xgb(forest, X, base_score, n_threads=1):
    """XGBoost gbtree/reg:squarederror semantics (fp32 sequential margin). Returns float32[n]."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    assert X.ndim == 2 and X.shape[1] == int(forest["n_features"])
    f, keep = _keep(forest)
    out = np.empty(X.shape[0], dtype=np.float32)
    lib().oracle_forest_predict_xgb(
        ctypes.byref(f), ctypes.c_void_p(X.ctypes.data), ctypes.c_int64(X.shape[0]),
        ctypes.c_float(float(np.float32(base_score))), ctypes.c_void_p(out.ctypes.data),
        ctypes.c_int(n_threads))
    del keep
    return out


def xgb_sigmoid(margin):
    """binary:logistic / reg:logistic link as the xgboost CPU library computes it (fp32 expf)."""
    y = np.ascontiguousarray(margin, dtype=np.float32).copy()
    lib().oracle_xgb_sigmoid(ctypes.c_void_p(y.ctypes.data), ctypes.c_int64(y.size))
    return y


def xgb_prob_to_margin(base_score):
    f = lib().oracle_xgb_prob_to_margin
    f.restype = ctypes.c_float
    f.argtypes = [ctypes.c_float]
    return float(f(ctypes.c_float(float(np.float32(base_score)))))


def forest_predict_f64(forest, X, init, scale, divisor, n_threads=1):
    """sklearn tree-ensemble semantics (fp64 sequential). Returns float64[n]."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    assert X.ndim == 2 and X.shape[1] == int(forest["n_features"])
    f, keep = _keep(forest)
    out = np.empty(X.shape[0], dtype=np.float64)
    lib().oracle_forest_predict_f64(
        ctypes.byref(f), ctypes.c_void_p(X.ctypes.data), ctypes.c_int64(X.shape[0]),
        ctypes.c_double(init), ctypes.c_double(scale), ctypes.c_double(divisor),
        ctypes.c_void_p(out.ctypes.data), ctypes.c_int(n_threads))
    del keep
    return out


def linear_predict(X, W, b):
    """fp64 decision function + labels (index of the winning class / (score>0))."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    W = np.ascontiguousarray(np.atleast_2d(W), dtype=np.float64)
    b = np.ascontiguousarray(np.atleast_1d(b), dtype=np.float64)
    n, k = X.shape
    c = W.shape[0]
    scores = np.empty((n, c), dtype=np.float64)
    labels = np.empty(n, dtype=np.int64)
    lib().oracle_linear_predict(
        ctypes.c_void_p(X.ctypes.data), ctypes.c_int64(n), ctypes.c_int32(k),
        ctypes.c_void_p(W.ctypes.data), ctypes.c_void_p(b.ctypes.data), ctypes.c_int32(c),
        ctypes.c_void_p(scores.ctypes.data), ctypes.c_void_p(labels.ctypes.data))
    return scores, labels


# ----------------------------------------------------------------------------------------------
# Source-form builders (shared by tests, gen_golden.py and bench.py)
# ----------------------------------------------------------------------------------------------

def forest_from_sklearn(estimators, n_features):
    """Concatenate sklearn DecisionTreeRegressor.tree_ arrays into source form.
    sklearn semantics: children_left == -1 marks a leaf; value[node,0,0] is the leaf output;
    missing_go_to_left (sklearn>=1.3) is the NaN direction."""
    off, left, right, feat, thr, dl, val = [0], [], [], [], [], [], []
    for e in estimators:
        t = e.tree_
        left.append(t.children_left.astype(np.int32))
        right.append(t.children_right.astype(np.int32))
        feat.append(np.maximum(t.feature, 0).astype(np.int32))
        thr.append(t.threshold.astype(np.float64))
        mg = getattr(t, "missing_go_to_left", None)
        dl.append(np.asarray(mg, dtype=np.uint8) if mg is not None
                  else np.zeros(t.node_count, np.uint8))
        val.append(t.value[:, 0, 0].astype(np.float64))
        off.append(off[-1] + t.node_count)
    return dict(tree_offset=np.asarray(off, np.int32), left=np.concatenate(left),
                right=np.concatenate(right), feat=np.concatenate(feat), thr=np.concatenate(thr),
                default_left=np.concatenate(dl), value=np.concatenate(val),
                n_features=int(n_features))


def synth_xgb_forest(n_trees=1000, depth=6, n_features=32, seed=0, ragged=False):
    """The BASELINE.json configs[1] model: complete depth-`depth` trees,
    split_indices~U{0..F-1}, split_conditions~N(0,1) fp32, leaves~N(0,0.1) fp32,
    default_left~Bernoulli(.5), base_score 0.5.  Node order inside a tree follows XGBoost's
    allocation (children allocated in pairs, breadth-first).  ragged=True prunes random
    subtrees so trees are incomplete (edge-case coverage)."""
    rng = np.random.default_rng(seed)
    off, left, right, feat, thr, dl, val = [0], [], [], [], [], [], []
    for _ in range(n_trees):
        # build breadth-first; each queue entry = (node id, depth)
        l, r, f, th, d, v = [], [], [], [], [], []
        nodes = [(0, 0)]
        l.append(-1); r.append(-1); f.append(0); th.append(0.0); d.append(0); v.append(0.0)
        qi = 0
        while qi < len(nodes):
            nid, dep = nodes[qi]
            qi += 1
            is_leaf = dep == depth or (ragged and dep > 0 and rng.random() < 0.25)
            if is_leaf:
                v[nid] = float(np.float32(rng.normal(0.0, 0.1)))
                continue
            f[nid] = int(rng.integers(0, n_features))
            th[nid] = float(np.float32(rng.normal(0.0, 1.0)))
            d[nid] = int(rng.random() < 0.5)
            li = len(l)
            for _k in range(2):
                l.append(-1); r.append(-1); f.append(0); th.append(0.0); d.append(0); v.append(0.0)
            l[nid], r[nid] = li, li + 1
            nodes.append((li, dep + 1))
            nodes.append((li + 1, dep + 1))
        left.append(np.asarray(l, np.int32)); right.append(np.asarray(r, np.int32))
        feat.append(np.asarray(f, np.int32)); thr.append(np.asarray(th, np.float64))
        dl.append(np.asarray(d, np.uint8)); val.append(np.asarray(v, np.float64))
        off.append(off[-1] + len(l))
    return dict(tree_offset=np.asarray(off, np.int32), left=np.concatenate(left),
                right=np.concatenate(right), feat=np.concatenate(feat), thr=np.concatenate(thr),
                default_left=np.concatenate(dl), value=np.concatenate(val),
                n_features=int(n_features))


def xgb_json_from_forest(forest, base_score=0.5, objective="reg:squarederror"):
    """Serialise a source-form forest in the XGBoost >=1.0 JSON model schema
    (learner.gradient_booster.model.trees[*]): what `Booster.save_model('x.json')` writes and
    what the reference loads with Booster.load_model (preprocess_service.py:475-476).
    In that schema a leaf stores its value in split_conditions and base_weights."""
    trees = []
    off = forest["tree_offset"]
    for t in range(len(off) - 1):
        s, e = int(off[t]), int(off[t + 1])
        left = forest["left"][s:e]
        is_leaf = left < 0
        cond = np.where(is_leaf, forest["value"][s:e], forest["thr"][s:e]).astype(np.float32)
        parents = np.full(e - s, 2147483647, dtype=np.int64)
        for i in range(e - s):
            if left[i] >= 0:
                parents[left[i]] = i
                parents[forest["right"][s + i]] = i
        trees.append(dict(
            base_weights=[float(x) for x in cond],
            categories=[], categories_nodes=[], categories_segments=[], categories_sizes=[],
            default_left=[int(x) for x in forest["default_left"][s:e]],
            id=t,
            left_children=[int(x) for x in left],
            loss_changes=[0.0] * (e - s),
            parents=[int(x) for x in parents],
            right_children=[int(x) for x in forest["right"][s:e]],
            split_conditions=[float(x) for x in cond],
            split_indices=[int(x) for x in forest["feat"][s:e]],
            split_type=[0] * (e - s),
            sum_hessian=[1.0] * (e - s),
            tree_param=dict(num_deleted="0", num_feature=str(forest["n_features"]),
                            num_nodes=str(e - s), size_leaf_vector="0"),
        ))
    n_trees = len(trees)
    return dict(
        learner=dict(
            attributes={},
            feature_names=[], feature_types=[],
            gradient_booster=dict(
                model=dict(
                    gbtree_model_param=dict(num_parallel_tree="1", num_trees=str(n_trees),
                                            size_leaf_vector="0"),
                    tree_info=[0] * n_trees,
                    trees=trees),
                name="gbtree"),
            learner_model_param=dict(base_score=repr(float(np.float32(base_score))),
                                     boost_from_average="1", num_class="0",
                                     num_feature=str(forest["n_features"]), num_target="1"),
            objective=dict(name=objective, reg_loss_param=dict(scale_pos_weight="1")),
        ),
        version=[1, 7, 6],
    )
