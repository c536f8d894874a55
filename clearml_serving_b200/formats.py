"""Model ingestion: turn the files the reference's engines load into packed blobs for libb200serve.

Reference loaders replaced (load-time only; nothing here is on the request path):
  * xgboost.Booster().load_model(path)   clearml_serving/serving/preprocess_service.py:475-476
  * joblib.load(path) of an sklearn model clearml_serving/serving/preprocess_service.py:457
The device layouts are documented in csrc/forest.cu and csrc/linear.cu.
"""
import json
import struct

import numpy as np

from . import native

_IDENTITY_OBJECTIVES = ("reg:squarederror", "reg:linear", "reg:absoluteerror", "reg:pseudohubererror")
# objectives whose base_score is a PROBABILITY: the booster adds logit(base_score) to the margin
# (xgboost LogisticRegression::ProbToMargin; LogisticRaw inherits it) ...
_LOGIT_BASE_OBJECTIVES = ("binary:logitraw", "binary:logistic", "reg:logistic")
# ... and those whose prediction is sigmoid(margin) (PredTransform)
_SIGMOID_OBJECTIVES = ("binary:logistic", "reg:logistic")
LINK_IDENTITY, LINK_SIGMOID = 0, 1


class PackedModel(object):
    """kind + blob + the I/O description the engine needs."""

    def __init__(self, kind, blob, description):
        self.kind = kind
        self.blob = blob
        self.description = description


# --------------------------------------------------------------------------------------------
# forests
# --------------------------------------------------------------------------------------------

def _skl_threshold_to_f32_strict(thr64):
    """sklearn goes left iff fp32 x <= thr64.  Exact fp32 restatement: x < nextafter(down32(thr64), +inf)
    where down32 rounds toward -inf to fp32 (for fp32 x: x <= thr64 <=> x <= down32(thr64))."""
    thr64 = np.asarray(thr64, dtype=np.float64)
    with np.errstate(over="ignore"):
        t = thr64.astype(np.float32)
    too_big = t.astype(np.float64) > thr64
    t = np.where(too_big, np.nextafter(t, np.float32(-np.inf)), t).astype(np.float32)
    return np.nextafter(t, np.float32(np.inf)).astype(np.float32)


def pack_forest(forest, mode, base=0.0, scale=1.0, divisor=1.0, link=LINK_IDENTITY):
    """forest: source form (dict of arrays: tree_offset,left,right,feat,thr,default_left,value,
    n_features; children local to each tree, left<0 => leaf).
    mode 'xgb': fp32 sequential sum, split `x < float32(thr)`.
    mode 'skl': fp64 sequential sum of scale*value, split `x <= thr64`, result / divisor.
    link: LINK_SIGMOID applies xgboost's fp32 logistic transform to the margin on the device ('xgb' mode only)."""
    if link not in (LINK_IDENTITY, LINK_SIGMOID) or (link != LINK_IDENTITY and mode != "xgb"):
        raise ValueError("pack_forest: bad link")
    if mode not in ("xgb", "skl"):
        raise ValueError("pack_forest: mode must be 'xgb' or 'skl'")
    off = np.asarray(forest["tree_offset"], dtype=np.int64)
    n_trees = len(off) - 1
    n_features = int(forest["n_features"])
    if n_trees <= 0:
        raise ValueError("pack_forest: empty forest")
    left = np.asarray(forest["left"], dtype=np.int64)
    right = np.asarray(forest["right"], dtype=np.int64)
    feat = np.asarray(forest["feat"], dtype=np.int64)
    thr = np.asarray(forest["thr"], dtype=np.float64)
    dl = np.asarray(forest["default_left"], dtype=np.int64)
    value = np.asarray(forest["value"], dtype=np.float64)

    feat_bits = max(1, int(n_features - 1).bit_length())
    left_bits = 31 - feat_bits
    f64 = mode == "skl"
    thr32 = _skl_threshold_to_f32_strict(thr) if f64 else thr.astype(np.float32)

    new_off = [0]
    val_bits = []
    meta = []
    leaf64 = []
    for t in range(n_trees):
        s, e = int(off[t]), int(off[t + 1])
        n = e - s
        if n <= 0:
            raise ValueError("pack_forest: tree {} has no nodes".format(t))
        if n >= (1 << left_bits):
            raise ValueError("pack_forest: tree {} has {} nodes; {} features leave room for {}".format(
                t, n, n_features, (1 << left_bits) - 1))
        # breadth-first renumbering so that right child == left child + 1
        order = [0]
        new_left = {}
        qi = 0
        while qi < len(order):
            old = order[qi]
            l_old = int(left[s + old])
            if l_old >= 0:
                r_old = int(right[s + old])
                if not (0 <= l_old < n and 0 <= r_old < n):
                    raise ValueError("pack_forest: tree {} has a child index out of range".format(t))
                new_left[qi] = len(order)
                order.append(l_old)
                order.append(r_old)
                if len(order) > n:
                    raise ValueError("pack_forest: tree {} is not a tree (node reached twice)".format(t))
            qi += 1
        order = np.asarray(order, dtype=np.int64) + s
        is_leaf = left[order] < 0
        m = np.zeros(len(order), dtype=np.uint32)
        v = np.zeros(len(order), dtype=np.uint32)
        for new_i, nl in new_left.items():
            g = order[new_i]
            f = int(feat[g])
            if not (0 <= f < n_features):
                raise ValueError("pack_forest: feature index {} out of range".format(f))
            m[new_i] = np.uint32(f | (int(dl[g] != 0) << feat_bits) | (nl << (feat_bits + 1)))
        v[~is_leaf] = thr32[order[~is_leaf]].view(np.uint32)
        if f64:
            lv = (np.float64(scale) * value[order[is_leaf]]).astype(np.float64)
            v[is_leaf] = np.arange(len(leaf64), len(leaf64) + lv.size, dtype=np.uint32)
            leaf64.extend(lv.tolist())
        else:
            v[is_leaf] = value[order[is_leaf]].astype(np.float32).view(np.uint32)
        val_bits.append(v)
        meta.append(m)
        new_off.append(new_off[-1] + len(order))

    val_bits = np.concatenate(val_bits)
    meta = np.concatenate(meta)
    nodes = (val_bits.astype(np.uint64) | (meta.astype(np.uint64) << np.uint64(32))).astype("<u8")
    toff = np.asarray(new_off, dtype="<u4")
    toff_bytes = toff.tobytes()
    if len(toff_bytes) % 8:
        toff_bytes += b"\0" * (8 - len(toff_bytes) % 8)
    leaf64 = np.asarray(leaf64, dtype="<f8")
    base_v = float(base) if f64 else float(np.float32(base))
    header = struct.pack("<4sIIIIIIIIIdd", b"B2SF", 1, n_trees, n_features, int(nodes.size), feat_bits,
                         1 if f64 else 0, int(leaf64.size), int(link), 0, base_v, float(divisor))
    blob = header + toff_bytes + nodes.tobytes() + leaf64.tobytes()
    desc = dict(kind="forest", mode=mode, n_trees=n_trees, n_features=n_features, n_nodes=int(nodes.size),
                input_dtype="float32", output_dtype="float64" if f64 else "float32", link=int(link))
    return PackedModel(native.MODEL_FOREST, blob, desc)


def _xgb_base_score(text):
    """learner_model_param.base_score: "5E-1", or the vector form "[5E-1]" newer releases write"""
    t = str(text).strip()
    if t.startswith("["):
        vals = [v for v in t.strip("[]").split(",") if v.strip()]
        if len(vals) != 1:
            raise ValueError("b200 engine: multi-target base_score {} is not supported".format(text))
        t = vals[0]
    return np.float32(float(t))


def _xgb_objective_plan(objective, base_score, base_is_margin=False):
    """objective name + stored base_score -> (base margin as fp32, link).  PARITY UNPINNED (xgboost not installable)."""
    if objective in _IDENTITY_OBJECTIVES:
        return np.float32(base_score), LINK_IDENTITY
    if objective in _LOGIT_BASE_OBJECTIVES:
        bs = np.float32(base_score)
        if base_is_margin:      # files written by xgboost < 1.0 store the margin itself
            margin = bs
        else:
            if not (0.0 < float(bs) < 1.0):
                raise ValueError("b200 engine: base_score {} must be in (0, 1) for objective '{}'".format(float(bs), objective))
            with np.errstate(all="ignore"):
                margin = np.float32(-np.log(np.float32(np.float32(1.0) / bs - np.float32(1.0))))   # -logf(1/bs - 1), fp32
        return np.float32(margin), LINK_SIGMOID if objective in _SIGMOID_OBJECTIVES else LINK_IDENTITY
    raise ValueError("b200 engine: XGBoost objective '{}' is not supported yet (supported: {})".format(
        objective, ", ".join(_IDENTITY_OBJECTIVES + _LOGIT_BASE_OBJECTIVES)))


def parse_xgboost_json(model):
    """XGBoost JSON model schema (learner.gradient_booster.model.trees[*]) -> (source form, base margin, link).
    Accepts a dict, a JSON string/bytes, or a path."""
    if isinstance(model, (bytes, bytearray)):
        model = json.loads(model.decode("utf-8"))
    elif isinstance(model, str):
        if model.lstrip().startswith("{"):
            model = json.loads(model)
        else:
            with open(model, "rb") as f:
                head = f.read(1)
                f.seek(0)
                if head != b"{":
                    raise ValueError(
                        "b200 engine: '{}' is not an XGBoost JSON model (save it with "
                        "Booster.save_model('model.json'))".format(model))
                model = json.load(f)
    learner = model["learner"]
    objective = learner.get("objective", {}).get("name", "reg:squarederror")
    lmp = learner["learner_model_param"]
    base_margin, link = _xgb_objective_plan(objective, _xgb_base_score(lmp.get("base_score", "0.5")))
    if int(lmp.get("num_class", "0") or 0) > 1 or int(lmp.get("num_target", "1") or 1) > 1:
        raise ValueError("b200 engine: multi-class / multi-target XGBoost models are not supported yet")
    gb = learner["gradient_booster"]
    if gb.get("name", "gbtree") not in ("gbtree",):
        raise ValueError("b200 engine: booster '{}' is not supported (gbtree only)".format(gb.get("name")))
    n_features = int(lmp["num_feature"])
    trees = gb["model"]["trees"]
    off, left, right, feat, thr, dl, val = [0], [], [], [], [], [], []
    for t in trees:
        if any(int(x) != 0 for x in t.get("split_type", [])) or len(t.get("categories", [])):
            raise ValueError("b200 engine: categorical splits are not supported")
        l = np.asarray(t["left_children"], dtype=np.int32)
        cond = np.asarray(t["split_conditions"], dtype=np.float64)
        left.append(l)
        right.append(np.asarray(t["right_children"], dtype=np.int32))
        feat.append(np.asarray(t["split_indices"], dtype=np.int32))
        thr.append(cond)
        dl.append(np.asarray(t["default_left"], dtype=np.uint8))
        val.append(cond)  # a leaf keeps its value in split_conditions
        off.append(off[-1] + len(l))
    forest = dict(tree_offset=np.asarray(off, np.int32), left=np.concatenate(left),
                  right=np.concatenate(right), feat=np.concatenate(feat), thr=np.concatenate(thr),
                  default_left=np.concatenate(dl), value=np.concatenate(val), n_features=n_features)
    return forest, float(base_margin), link


def pack_xgboost_json(model):
    forest, base_margin, link = parse_xgboost_json(model)
    return pack_forest(forest, "xgb", base=base_margin, link=link)


# --------------------------------------------------------------------------------------------
# XGBoost legacy binary models: what `Booster.save_model("xgb_model")` (no .json / .ubj extension) writes -- the
# reference's own example does exactly that (examples/xgboost/train_model.py:28) and its engine loads the file
# with Booster.load_model (preprocess_service.py:475-476).  Layout restated from xgboost 1.7 (src/learner.cc
# LearnerModelParamLegacy / LearnerIO::LoadModel, src/gbm/gbtree_model.h GBTreeModelParam, include/xgboost/tree_model.h
# TreeParam / RegTree::Node / RTreeNodeStat; little endian):
#   ["binf"]                                        optional 4-byte tag of pre-1.0 writers
#   LearnerModelParamLegacy  136 B                  f32 base_score | u32 num_feature | i32 num_class |
#                                                   i32 contain_extra_attrs | i32 contain_eval_metrics |
#                                                   u32 major_version | u32 minor_version | u32 num_target | reserved
#   u64 len + bytes          objective name         ("reg:squarederror")
#   u64 len + bytes          booster name           ("gbtree")
#   GBTreeModelParam         160 B                  i32 num_trees | i32 num_parallel_tree | i32 | i32 pad |
#                                                   i64 num_pbuffer | i32 num_output_group | i32 size_leaf_vector | reserved
#   per tree: TreeParam 148 B (i32 num_roots | i32 num_nodes | i32 num_deleted | i32 max_depth | i32 num_feature |
#             i32 size_leaf_vector | reserved), num_nodes x Node 20 B (i32 parent | i32 cleft | i32 cright |
#             u32 sindex (bit 31 = default left) | f32 leaf value / split condition), num_nodes x RTreeNodeStat 16 B
#   num_trees x i32 tree_info; then optional attributes / metric names (ignored here)
# xgboost is not installable here and the reference ships no model file: PARITY UNPINNED (tests/test_formats.py
# round-trips a writer of the same published layout).
# --------------------------------------------------------------------------------------------
def looks_like_xgboost_legacy_binary(head):
    """`head`: the first >= 160 bytes of a file"""
    o = 4 if head[:4] == b"binf" else 0
    if len(head) < o + 136 + 8 + 4:
        return False
    n = struct.unpack_from("<Q", head, o + 136)[0]
    name = head[o + 144:o + 144 + min(n, 32)]
    return 3 <= n <= 64 and b":" in name and all(32 <= c < 127 for c in name)


def parse_xgboost_legacy_binary(data):
    """bytes of a legacy binary model -> (source form, base margin, link)"""
    data = bytes(data)
    o = 4 if data[:4] == b"binf" else 0
    if data[:4] == b"bs64":
        raise ValueError("b200 engine: base64 XGBoost models (xgboost < 0.6) are not supported")

    def need(n):
        if o + n > len(data):
            raise ValueError("b200 engine: truncated XGBoost binary model")
    need(136)
    base_score, num_feature, num_class, has_attrs, has_metrics, major, _minor, num_target = struct.unpack_from("<fIiiiIII", data, o)
    o += 136

    def read_str():
        nonlocal o
        need(8)
        n = struct.unpack_from("<Q", data, o)[0]
        o += 8
        if n > 256:
            raise ValueError("b200 engine: not an XGBoost binary model (implausible string length)")
        need(n)
        sv = data[o:o + n].decode("ascii", "replace")
        o += n
        return sv
    objective = read_str()
    booster = read_str()
    if booster != "gbtree":
        raise ValueError("b200 engine: booster '{}' is not supported (gbtree only)".format(booster))
    if num_class > 1 or num_target > 1:
        raise ValueError("b200 engine: multi-class / multi-target XGBoost models are not supported yet")
    need(160)
    num_trees, _npt, _nf, _pad, num_pbuffer, _nog, size_leaf_vector = struct.unpack_from("<iiiiqii", data, o)
    o += 160
    if num_trees <= 0 or size_leaf_vector not in (0, 1):
        raise ValueError("b200 engine: XGBoost binary model with {} trees / leaf vectors of {}".format(num_trees, size_leaf_vector))
    off, left, right, feat, thr, dl, val = [0], [], [], [], [], [], []
    node_t = np.dtype([("parent", "<i4"), ("cleft", "<i4"), ("cright", "<i4"), ("sindex", "<u4"), ("info", "<f4")])
    for _t in range(num_trees):
        need(148)
        _roots, num_nodes, _deleted, _depth, _tnf, _slv = struct.unpack_from("<iiiiii", data, o)
        o += 148
        if num_nodes <= 0:
            raise ValueError("b200 engine: XGBoost binary model: tree without nodes")
        need(num_nodes * 36)
        nodes = np.frombuffer(data, dtype=node_t, count=num_nodes, offset=o)
        o += num_nodes * 36          # nodes (20 B each) + statistics (16 B each)
        l = nodes["cleft"].astype(np.int32)
        left.append(l)
        right.append(np.where(l < 0, -1, nodes["cright"]).astype(np.int32))
        feat.append(np.where(l < 0, 0, nodes["sindex"] & np.uint32(0x7FFFFFFF)).astype(np.int32))
        dl.append(((nodes["sindex"] >> np.uint32(31)) & np.uint32(1)).astype(np.uint8))
        cond = nodes["info"].astype(np.float64)
        thr.append(cond)
        val.append(cond)
        off.append(off[-1] + num_nodes)
    need(num_trees * 4)
    o += num_trees * 4               # tree_info (output group of every tree: all 0 for single-output models)
    forest = dict(tree_offset=np.asarray(off, np.int32), left=np.concatenate(left), right=np.concatenate(right),
                  feat=np.concatenate(feat), thr=np.concatenate(thr), default_left=np.concatenate(dl),
                  value=np.concatenate(val), n_features=int(num_feature))
    if forest["feat"].size and int(forest["feat"].max()) >= int(num_feature):
        raise ValueError("b200 engine: XGBoost binary model: split feature out of range")
    base_margin, link = _xgb_objective_plan(objective, base_score, base_is_margin=major < 1)
    return forest, float(base_margin), link


def pack_xgboost_legacy_binary(data):
    if isinstance(data, str):
        with open(data, "rb") as f:
            data = f.read()
    forest, base_margin, link = parse_xgboost_legacy_binary(data)
    return pack_forest(forest, "xgb", base=base_margin, link=link)


# --------------------------------------------------------------------------------------------
# LightGBM text models (`Booster.save_model("model.txt")`, what `LightGBMPreprocessRequest` loads with
# `lgbm.Booster(model_file=...)`, clearml_serving/serving/preprocess_service.py:486-501; lightgbm 3.3.x,
# requirements.txt:17).  lightgbm is not installable here and the reference has no golden vectors for it: the semantics are
# restated from the library's published model format and predictor (Tree::Predict: numerical decision `fval <= threshold`
# on double; child index >= 0 = internal node, < 0 = leaf ~index; decision_type bit 1 = default left, bits 2-3 = missing
# type {0 none: NaN reads as 0.0, 1 zero, 2 NaN: missing goes to the default side}; the raw score is the sum of
# `leaf_value` over trees in order, in double; `average_output` (boosting=rf) divides by the number of trees).
# PARITY UNPINNED like the xgboost row of DESIGN.md section 3.
# --------------------------------------------------------------------------------------------
_LGB_IDENTITY = ("regression", "regression_l1", "huber", "fair", "quantile", "mape", "custom", "none")


def parse_lightgbm_text(model):
    """LightGBM model text (or a path to it) -> (source form for pack_forest, divisor)"""
    if isinstance(model, (bytes, bytearray)):
        model = model.decode("utf-8")
    if "\n" not in model:
        with open(model, "rt") as f:
            model = f.read()
    head, _, rest = model.partition("Tree=")
    hdr = dict(ln.split("=", 1) for ln in head.splitlines() if "=" in ln)
    if not head.lstrip().startswith("tree") or "max_feature_idx" not in hdr:
        raise ValueError("b200 engine: not a LightGBM text model (save it with Booster.save_model('model.txt'))")
    objective = hdr.get("objective", "regression").split()[0]
    if objective not in _LGB_IDENTITY:
        raise ValueError("b200 engine: LightGBM objective '{}' is not supported yet (identity-link objectives only: {})".format(
            objective, ", ".join(_LGB_IDENTITY)))
    if int(hdr.get("num_class", "1")) != 1 or int(hdr.get("num_tree_per_iteration", "1")) != 1:
        raise ValueError("b200 engine: multi-class LightGBM models are not supported yet")
    n_features = int(hdr["max_feature_idx"]) + 1
    average = "average_output" in head.split()
    off, left, right, feat, thr, dl, val = [0], [], [], [], [], [], []
    body = ("Tree=" + rest).split("end of trees")[0]
    for block in body.split("Tree=")[1:]:
        kv = dict(ln.split("=", 1) for ln in block.splitlines() if "=" in ln)
        n_leaves = int(kv["num_leaves"])
        if int(kv.get("num_cat", "0")) != 0:
            raise ValueError("b200 engine: categorical splits are not supported")
        if int(kv.get("is_linear", "0")) != 0:
            raise ValueError("b200 engine: LightGBM linear trees are not supported")
        leaf_value = np.array(kv["leaf_value"].split(), dtype=np.float64)
        n_int = n_leaves - 1
        if leaf_value.size != n_leaves:
            raise ValueError("b200 engine: LightGBM tree with {} leaf values for {} leaves".format(leaf_value.size, n_leaves))
        if n_int == 0:   # a stump that never split: one leaf
            left.append(np.array([-1], np.int32)); right.append(np.array([-1], np.int32)); feat.append(np.zeros(1, np.int32))
            thr.append(np.zeros(1)); dl.append(np.zeros(1, np.uint8)); val.append(leaf_value[:1])
            off.append(off[-1] + 1)
            continue
        sf = np.array(kv["split_feature"].split(), dtype=np.int64)
        th = np.array(kv["threshold"].split(), dtype=np.float64)
        dt = np.array(kv["decision_type"].split(), dtype=np.int64)
        lc = np.array(kv["left_child"].split(), dtype=np.int64)
        rc = np.array(kv["right_child"].split(), dtype=np.int64)
        if not (sf.size == th.size == dt.size == lc.size == rc.size == n_int):
            raise ValueError("b200 engine: LightGBM tree arrays disagree on the number of internal nodes")
        if np.any(dt & 1):
            raise ValueError("b200 engine: categorical splits are not supported")
        missing = (dt >> 2) & 3
        if np.any(missing == 1):
            raise ValueError("b200 engine: LightGBM zero_as_missing splits are not supported")
        if sf.max() >= n_features or sf.min() < 0:
            raise ValueError("b200 engine: LightGBM split feature out of range")
        # missing type 2 (NaN): the stored default side; type 0 (none): NaN reads as 0.0, so its side is known now
        default_left = np.where(missing == 2, (dt >> 1) & 1, (0.0 <= th).astype(np.int64))
        child = lambda c: np.where(c >= 0, c, n_int + (~c))   # noqa: E731 -- leaves are appended after the internal nodes
        n = n_int + n_leaves
        l_all, r_all = np.full(n, -1, np.int32), np.full(n, -1, np.int32)
        l_all[:n_int], r_all[:n_int] = child(lc), child(rc)
        if l_all[:n_int].max() >= n or r_all[:n_int].max() >= n or l_all[:n_int].min() < 0 or r_all[:n_int].min() < 0:
            raise ValueError("b200 engine: LightGBM child index out of range")
        f_all, t_all, d_all, v_all = np.zeros(n, np.int32), np.zeros(n), np.zeros(n, np.uint8), np.zeros(n)
        f_all[:n_int], t_all[:n_int], d_all[:n_int] = sf, th, default_left
        v_all[n_int:] = leaf_value
        left.append(l_all); right.append(r_all); feat.append(f_all); thr.append(t_all); dl.append(d_all); val.append(v_all)
        off.append(off[-1] + n)
    if len(off) == 1:
        raise ValueError("b200 engine: LightGBM model without trees")
    forest = dict(tree_offset=np.asarray(off, np.int32), left=np.concatenate(left), right=np.concatenate(right),
                  feat=np.concatenate(feat), thr=np.concatenate(thr), default_left=np.concatenate(dl),
                  value=np.concatenate(val), n_features=n_features)
    return forest, float(len(off) - 1) if average else 1.0


def pack_lightgbm_text(model):
    forest, divisor = parse_lightgbm_text(model)
    return pack_forest(forest, "skl", base=0.0, scale=1.0, divisor=divisor)


def _forest_from_sklearn_trees(estimators, n_features):
    off, left, right, feat, thr, dl, val = [0], [], [], [], [], [], []
    for e in estimators:
        t = e.tree_
        if t.value.shape[1] != 1 or t.value.shape[2] != 1:
            raise ValueError("b200 engine: multi-output / classification trees are not supported yet")
        left.append(t.children_left.astype(np.int32))
        right.append(t.children_right.astype(np.int32))
        feat.append(np.maximum(t.feature, 0).astype(np.int32))
        thr.append(t.threshold.astype(np.float64))
        mg = getattr(t, "missing_go_to_left", None)
        dl.append(np.asarray(mg, dtype=np.uint8) if mg is not None else np.zeros(t.node_count, np.uint8))
        val.append(t.value[:, 0, 0].astype(np.float64))
        off.append(off[-1] + t.node_count)
    return dict(tree_offset=np.asarray(off, np.int32), left=np.concatenate(left), right=np.concatenate(right),
                feat=np.concatenate(feat), thr=np.concatenate(thr), default_left=np.concatenate(dl),
                value=np.concatenate(val), n_features=int(n_features))


# --------------------------------------------------------------------------------------------
# linear
# --------------------------------------------------------------------------------------------

def pack_linear(coef, intercept, classes):
    coef = np.atleast_2d(np.asarray(coef, dtype=np.float64))
    intercept = np.atleast_1d(np.asarray(intercept, dtype=np.float64))
    classes = np.asarray(classes, dtype=np.int64)
    n_out, n_features = coef.shape
    if intercept.size != n_out:
        raise ValueError("pack_linear: intercept size mismatch")
    if classes.size != (2 if n_out == 1 else n_out):
        raise ValueError("pack_linear: classes size mismatch")
    header = struct.pack("<4sIIIII", b"B2SL", 1, n_features, n_out, int(classes.size), 0)
    blob = header + coef.astype("<f8").tobytes() + intercept.astype("<f8").tobytes() + classes.astype("<i8").tobytes()
    desc = dict(kind="linear", n_features=n_features, n_out=n_out, input_dtype="float64", output_dtype="int64")
    return PackedModel(native.MODEL_LINEAR, blob, desc)


# --------------------------------------------------------------------------------------------
# dispatch on what the reference's engines would have loaded
# --------------------------------------------------------------------------------------------

def pack_sklearn(model):
    """An (unpickled) sklearn estimator -> PackedModel. Supported: GradientBoostingRegressor
    (squared_error), RandomForestRegressor / ExtraTreesRegressor, DecisionTreeRegressor,
    LogisticRegression-like linear classifiers (coef_/intercept_/classes_ with integer classes)."""
    name = type(model).__name__
    if name == "GradientBoostingRegressor":
        if getattr(model, "loss", "squared_error") != "squared_error":
            raise ValueError("b200 engine: GradientBoostingRegressor loss '{}' not supported".format(model.loss))
        init = getattr(model, "init_", None)
        const = getattr(init, "constant_", None)
        if const is None:
            raise ValueError("b200 engine: GradientBoostingRegressor with a custom init estimator is not supported")
        forest = _forest_from_sklearn_trees([e[0] for e in model.estimators_], model.n_features_in_)
        return pack_forest(forest, "skl", base=float(np.ravel(const)[0]), scale=float(model.learning_rate), divisor=1.0)
    if name in ("RandomForestRegressor", "ExtraTreesRegressor"):
        if getattr(model, "n_outputs_", 1) != 1:
            raise ValueError("b200 engine: multi-output forests are not supported")
        forest = _forest_from_sklearn_trees(model.estimators_, model.n_features_in_)
        return pack_forest(forest, "skl", base=0.0, scale=1.0, divisor=float(len(model.estimators_)))
    if name in ("DecisionTreeRegressor", "ExtraTreeRegressor"):
        forest = _forest_from_sklearn_trees([model], model.n_features_in_)
        # DecisionTreeRegressor.predict returns value[leaf]: 0 + v is exact, / 1 is exact
        return pack_forest(forest, "skl", base=0.0, scale=1.0, divisor=1.0)
    if hasattr(model, "coef_") and hasattr(model, "intercept_") and hasattr(model, "classes_"):
        classes = np.asarray(model.classes_)
        if not np.issubdtype(classes.dtype, np.integer):
            raise ValueError("b200 engine: only integer class labels are supported for linear classifiers")
        return pack_linear(model.coef_, model.intercept_, classes)
    raise ValueError("b200 engine: unsupported sklearn model type '{}'".format(name))


def pack_torch_module(module):
    """A torch module the reference would have exported for Triton (TorchScript / ONNX): lowered here
    instead.  Supported architectures: transformers BERT sequence classifiers."""
    arch = type(module).__name__
    if arch == "BertForSequenceClassification":
        return pack_bert(module)
    if arch == "ResNet":
        return pack_resnet(module.eval())
    raise ValueError("b200 engine: torch architecture '{}' is not supported yet".format(arch))


def load_model_file(path):
    """Sniff the model file the reference would hand to xgboost / joblib / Triton and pack it."""
    import os
    if os.path.isdir(path) and os.path.exists(os.path.join(path, "config.json")):
        # a transformers `save_pretrained` directory
        from transformers import AutoModelForSequenceClassification
        return pack_torch_module(AutoModelForSequenceClassification.from_pretrained(path).eval())
    with open(path, "rb") as f:
        head = f.read(16)
    if head.lstrip()[:1] == b"{":
        return pack_xgboost_json(path)
    import joblib  # noqa  (the reference's sklearn engine loads with joblib too)
    return pack_sklearn(joblib.load(path))


# --------------------------------------------------------------------------------------------
# DL graphs (B2S_MODEL_GRAPH): offline lowering of a torch module to the op list of csrc/graph.cu
# --------------------------------------------------------------------------------------------

OP_EMBED_LN, OP_LINEAR, OP_LAYERNORM, OP_ATTENTION, OP_GATHER_FIRST = 1, 2, 3, 4, 5
ACT_NONE, ACT_GELU, ACT_RELU, ACT_TANH = 0, 1, 2, 3
_DT = {"float32": 0, "float64": 1, "int32": 2, "int64": 3, "uint8": 4, "float16": 8}


class GraphBuilder(object):
    """Accumulates weights, activation buffers and ops, then serialises the "B2SG" blob."""

    def __init__(self, in_dtypes, max_pos, in_row_elems=None):
        self.tensors, self.buffers, self.ops = [], [], []
        self.in_dtypes = list(in_dtypes)
        self.in_row_elems = list(in_row_elems) if in_row_elems is not None else [-1] * len(self.in_dtypes)
        self.max_pos = int(max_pos)
        self.outputs = []

    def tensor(self, array, dtype):
        a = np.ascontiguousarray(np.asarray(array, dtype=dtype))
        self.tensors.append(a)
        return len(self.tensors) - 1

    def buffer(self, dtype, per_token, cols, rows_per_item=1):
        """per_token: one row per packed token (ragged models); else `rows_per_item` rows per batch item"""
        self.buffers.append((_DT[dtype], 0 if per_token else int(rows_per_item), int(cols)))
        return len(self.buffers) - 1

    def op(self, opcode, ints, floats=()):
        a = list(ints) + [0] * (15 - len(ints))
        f = list(floats) + [0.0] * (4 - len(floats))
        self.ops.append((opcode, a, f))

    def linear(self, in_buf, weight, bias, out_buf, act=ACT_NONE, residual=-1, out_f32=False, act_after=False):
        w = np.asarray(weight)
        n, k = w.shape
        wi = self.tensor(w, np.float16)
        bi = self.tensor(bias, np.float32) if bias is not None else -1
        self.op(OP_LINEAR, [in_buf, wi, bi, residual, out_buf, act, n, k, 1 if out_f32 else 0, 1 if act_after else 0])

    def output(self, buf):
        """marks `buf` as the next model output and returns the operand code ops use to write it"""
        self.outputs.append(buf)
        return -len(self.outputs)

    def serialise(self):
        n_t, n_b, n_o = len(self.tensors), len(self.buffers), len(self.ops)
        out_buf = (self.outputs + [0, 0, 0, 0])[:4]
        in_dt = ([_DT[d] for d in self.in_dtypes] + [0, 0, 0, 0])[:4]
        in_re = (list(self.in_row_elems) + [0, 0, 0, 0])[:4]
        header = struct.pack("<4s7I4i4i4q", b"B2SG", 1, n_t, n_b, n_o, len(self.in_dtypes), len(self.outputs),
                             self.max_pos, *out_buf, *in_dt, *in_re)
        assert len(header) == 96
        offsets, off = [], 0
        for a in self.tensors:
            offsets.append(off)
            off += (a.nbytes + 255) // 256 * 256
        tbl = b""
        for a, o in zip(self.tensors, offsets):
            shape = list(a.shape) + [1] * (4 - a.ndim)
            tbl += struct.pack("<II4qQQ", _DT[str(a.dtype)], a.ndim, *shape, o, a.nbytes)
        for (dt, kind, cols) in self.buffers:
            tbl += struct.pack("<IIq", dt, kind, cols)
        for (opc, a, f) in self.ops:
            tbl += struct.pack("<I15i4f", opc, *a, *f)
        head = header + tbl
        pad = (-len(head)) % 256
        data = bytearray(off)
        for a, o in zip(self.tensors, offsets):
            data[o:o + a.nbytes] = a.tobytes()
        return head + b"\0" * pad + bytes(data)


def pack_bert(model):
    """transformers BertForSequenceClassification (or BertModel + classifier-less) -> PackedModel.
    Inputs follow the reference's HF example (examples/huggingface/readme.md:113): int32
    input_ids, token_type_ids, attention_mask, one variable-length row per sequence; output fp32
    logits [batch, num_labels].  fp16 weights, fp32 biases / LayerNorm / residual stream."""
    cfg = model.config
    sd = {k: v.detach().cpu().float().numpy() for k, v in model.state_dict().items()}
    prefix = "bert." if any(k.startswith("bert.") for k in sd) else ""
    H, L, heads, I = cfg.hidden_size, cfg.num_hidden_layers, cfg.num_attention_heads, cfg.intermediate_size
    if H % heads or H // heads != 64:
        raise ValueError("b200 engine: attention head_dim {} not supported (64 only)".format(H // heads))
    if getattr(cfg, "hidden_act", "gelu") != "gelu":
        raise ValueError("b200 engine: hidden_act '{}' not supported (gelu only)".format(cfg.hidden_act))
    if getattr(cfg, "position_embedding_type", "absolute") != "absolute":
        raise ValueError("b200 engine: only absolute position embeddings are supported")
    eps = float(cfg.layer_norm_eps)
    g = GraphBuilder(["int32", "int32", "int32"], cfg.max_position_embeddings)
    h16, h32 = g.buffer("float16", True, H), g.buffer("float32", True, H)
    h16b, h32b = g.buffer("float16", True, H), g.buffer("float32", True, H)
    qkv, ctx = g.buffer("float16", True, 3 * H), g.buffer("float16", True, H)
    pre32, ff = g.buffer("float32", True, H), g.buffer("float16", True, I)
    e = prefix + "embeddings."
    g.op(OP_EMBED_LN, [0, 1, g.tensor(sd[e + "word_embeddings.weight"], np.float16),
                       g.tensor(sd[e + "position_embeddings.weight"], np.float16),
                       g.tensor(sd[e + "token_type_embeddings.weight"], np.float16),
                       g.tensor(sd[e + "LayerNorm.weight"], np.float32), g.tensor(sd[e + "LayerNorm.bias"], np.float32),
                       h16, h32, H, cfg.vocab_size, cfg.max_position_embeddings, cfg.type_vocab_size], [eps])
    for l in range(L):
        p = "{}encoder.layer.{}.".format(prefix, l)
        wqkv = np.concatenate([sd[p + "attention.self.query.weight"], sd[p + "attention.self.key.weight"],
                               sd[p + "attention.self.value.weight"]], 0)
        bqkv = np.concatenate([sd[p + "attention.self.query.bias"], sd[p + "attention.self.key.bias"],
                               sd[p + "attention.self.value.bias"]], 0)
        g.linear(h16, wqkv, bqkv, qkv)
        g.op(OP_ATTENTION, [qkv, 2, ctx, heads, 64])
        g.linear(ctx, sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"], pre32,
                 residual=h32, out_f32=True)
        g.op(OP_LAYERNORM, [pre32, g.tensor(sd[p + "attention.output.LayerNorm.weight"], np.float32),
                            g.tensor(sd[p + "attention.output.LayerNorm.bias"], np.float32), h16b, h32b, H], [eps])
        g.linear(h16b, sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"], ff, act=ACT_GELU)
        g.linear(ff, sd[p + "output.dense.weight"], sd[p + "output.dense.bias"], pre32, residual=h32b, out_f32=True)
        g.op(OP_LAYERNORM, [pre32, g.tensor(sd[p + "output.LayerNorm.weight"], np.float32),
                            g.tensor(sd[p + "output.LayerNorm.bias"], np.float32), h16, h32, H], [eps])
    cls16, pool16 = g.buffer("float16", False, H), g.buffer("float16", False, H)
    g.op(OP_GATHER_FIRST, [h16, cls16, H])
    g.linear(cls16, sd[prefix + "pooler.dense.weight"], sd[prefix + "pooler.dense.bias"], pool16, act=ACT_TANH)
    n_labels = sd["classifier.weight"].shape[0]
    logits = g.buffer("float32", False, n_labels)
    g.linear(pool16, sd["classifier.weight"], sd["classifier.bias"], g.output(logits), out_f32=True)
    flops_per_token = 2.0 * L * (4 * H * H + 2 * H * I)
    desc = dict(kind="graph", arch="bert", layers=L, hidden=H, heads=heads, intermediate=I, num_labels=int(n_labels),
                max_row_elems=int(cfg.max_position_embeddings), input_dtype="int32", output_dtype="float32",
                gemm_flops_per_token=flops_per_token, attn_flops_per_token_per_key=4.0 * L * H)
    return PackedModel(native.MODEL_GRAPH, g.serialise(), desc)


OP_NCHW_TO_NHWC, OP_IM2COL, OP_MAXPOOL, OP_AVGPOOL, OP_CONV = 6, 7, 8, 9, 10
OP_STEM_S2D, OP_CONV_STEM = 11, 12


def stem_s2d_weight(w):
    """7x7 filter [Cout, C<=4, 7, 7] -> [Cout, 256]: the stride-2 convolution rewritten as a 4x4 stride-1 convolution
    over the 2x2 space-to-depth image (csrc/conv.cu nchw_to_s2d_kernel): k = a*64 + b*16 + (dy*2 + dx)*4 + c holds
    w[co, c, 2a + dy, 2b + dx]; the taps with 2a + dy == 7 or 2b + dx == 7 and channels >= C are zero."""
    w = np.asarray(w, np.float64)
    Cout, C, KH, KW = w.shape
    if KH != 7 or KW != 7 or C > 4:
        raise ValueError("b200 engine: stem_s2d_weight wants a [Cout, <=4, 7, 7] filter")
    w8 = np.zeros((Cout, 4, 8, 8), np.float64)
    w8[:, :C, :7, :7] = w
    # [co, c, a, dy, b, dx] -> [co, a, b, dy, dx, c]
    return w8.reshape(Cout, 4, 4, 2, 4, 2).transpose(0, 2, 4, 3, 5, 1).reshape(Cout, 256)


def _fold_bn(conv_w, bn):
    """conv (no bias) followed by eval-mode BatchNorm -> (weight', bias')"""
    gamma = bn.weight.detach().double().numpy()
    beta = bn.bias.detach().double().numpy()
    mean = bn.running_mean.detach().double().numpy()
    var = bn.running_var.detach().double().numpy()
    scale = gamma / np.sqrt(var + bn.eps)
    w = conv_w.detach().double().numpy() * scale[:, None, None, None]
    return w, beta - mean * scale


class _ConvNetLowering(object):
    """Lowers torchvision-style conv / bn / relu / residual stacks onto GraphBuilder ops: NHWC fp16
    activations (channels padded to 8), 1x1 stride-1 convs as direct GEMMs, everything else as
    im2col + GEMM, BatchNorm folded, bias/ReLU/residual fused in the GEMM epilogue."""

    def __init__(self, g, implicit=True):
        self.g = g
        self.implicit = implicit   # False: every KxK / strided conv as explicit im2col + GEMM (comparison form)
        self._pool = {}   # (rows_per_item, cols, tag) -> [buffers], round robin

    def buf(self, rows_per_item, cols, tag="act", n=3):
        key = (int(rows_per_item), int(cols), tag)
        ring = self._pool.setdefault(key, dict(items=[], nxt=0))
        if len(ring["items"]) < n:
            ring["items"].append(self.g.buffer("float16", False, cols, rows_per_item=rows_per_item))
            return ring["items"][-1]
        b = ring["items"][ring["nxt"] % n]
        ring["nxt"] += 1
        return b

    def conv_bn(self, x, H, W, Cin_p, conv, bn, relu=True, residual=-1):
        g = self.g
        KH, KW = conv.kernel_size
        s, p = conv.stride[0], conv.padding[0]
        if conv.stride[0] != conv.stride[1] or conv.padding[0] != conv.padding[1] or conv.groups != 1 or conv.dilation != (1, 1):
            raise ValueError("b200 engine: unsupported convolution configuration")
        w, b = _fold_bn(conv.weight, bn) if bn is not None else (conv.weight.detach().double().numpy(),
                                                                 np.zeros(conv.out_channels))
        if conv.bias is not None:
            b = b + conv.bias.detach().double().numpy()
        Cout, Cin = w.shape[0], w.shape[1]
        if Cout % 8:
            raise ValueError("b200 engine: conv output channels must be a multiple of 8")
        OH, OW = (H + 2 * p - KH) // s + 1, (W + 2 * p - KW) // s + 1
        wk = np.zeros((Cout, KH, KW, Cin_p), np.float64)
        wk[:, :, :, :Cin] = np.transpose(w, (0, 2, 3, 1))
        K = KH * KW * Cin_p
        wk = wk.reshape(Cout, K)
        direct = KH == 1 and KW == 1 and s == 1 and p == 0
        # implicit GEMM (im2col-mode TMA) needs whole 64-channel k-blocks per filter tap; the 3-channel stem does not
        implicit = self.implicit and not direct and KH == KW and Cin_p % 64 == 0
        if direct:
            a = x
        elif not implicit:
            a = self.buf(OH * OW, K, tag="col", n=1)
            g.op(OP_IM2COL, [x, a, H, W, Cin_p, KH, KW, s, p, OH, OW, K])
        y = self.buf(OH * OW, Cout)
        guard = 0
        while y in (x, residual) and guard < 4:   # never write over a live operand
            y = self.buf(OH * OW, Cout)
            guard += 1
        act = ACT_RELU if relu else ACT_NONE
        if implicit:
            wi = g.tensor(wk, np.float16)
            bi = g.tensor(b, np.float32)
            g.op(OP_CONV, [x, wi, bi, residual, y, act, Cout, K, H, W, Cin_p, KH, s, p, 1 if residual >= 0 else 0])
        else:
            g.linear(a, wk, b, y, act=act, residual=residual, act_after=residual >= 0)
        return y, OH, OW, Cout


def pack_resnet(model, input_dtype="float32", image_hw=(224, 224), implicit_conv=None):
    """torchvision.models.resnet.ResNet (Bottleneck or BasicBlock) in eval mode -> PackedModel.
    Input: NCHW images [batch, 3, H, W] float32 (what the reference's Triton client can send,
    SURVEY.md F5) or uint8; output fp32 logits [batch, num_classes]."""
    H, W = image_hw
    g = GraphBuilder([input_dtype], 0, in_row_elems=[3 * H * W])
    if implicit_conv is None:
        import os
        implicit_conv = os.environ.get("B2S_CONV_IMPLICIT", "1") != "0"
    low = _ConvNetLowering(g, implicit=implicit_conv)
    c1 = model.conv1
    sOH, sOW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    stem_s2d = (implicit_conv and tuple(c1.kernel_size) == (7, 7) and tuple(c1.stride) == (2, 2) and
                tuple(c1.padding) == (3, 3) and c1.in_channels <= 4 and c1.groups == 1 and tuple(c1.dilation) == (1, 1) and
                c1.out_channels <= 128 and c1.out_channels % 8 == 0 and sOW <= 128)
    if stem_s2d:
        # the stem as a 4x4 stride-1 implicit GEMM over the 2x2 space-to-depth image (no patch matrix, K = 256)
        Hz, Wz = sOH + 3, sOW + 3
        z = g.buffer("float16", False, 16, rows_per_item=Hz * Wz)
        g.op(OP_STEM_S2D, [0, z, c1.in_channels, H, W, Hz, Wz])
        w, b = _fold_bn(c1.weight, model.bn1)
        if c1.bias is not None:
            b = b + c1.bias.detach().double().numpy()
        C = c1.out_channels
        x = low.buf(sOH * sOW, C)
        g.op(OP_CONV_STEM, [z, g.tensor(stem_s2d_weight(w), np.float16), g.tensor(b, np.float32), -1, x, ACT_RELU, C, 256,
                            Hz, Wz, sOH, sOW])
        H, W = sOH, sOW
    else:
        x = g.buffer("float16", False, 8, rows_per_item=H * W)
        g.op(OP_NCHW_TO_NHWC, [0, x, 3, H, W, 8])
        x, H, W, C = low.conv_bn(x, H, W, 8, model.conv1, model.bn1, relu=True)
    OH, OW = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    y = low.buf(OH * OW, C)
    g.op(OP_MAXPOOL, [x, y, H, W, C, OH, OW])
    x, H, W = y, OH, OW
    n_convs = 1
    for layer in (model.layer1, model.layer2, model.layer3, model.layer4):
        for blk in layer:
            kind = type(blk).__name__
            if blk.downsample is not None:
                idn, _, _, _ = low.conv_bn(x, H, W, C, blk.downsample[0], blk.downsample[1], relu=False)
                n_convs += 1
            else:
                idn = x
            if kind == "Bottleneck":
                t, h1, w1, c1 = low.conv_bn(x, H, W, C, blk.conv1, blk.bn1, relu=True)
                t, h2, w2, c2 = low.conv_bn(t, h1, w1, c1, blk.conv2, blk.bn2, relu=True)
                x, H, W, C = low.conv_bn(t, h2, w2, c2, blk.conv3, blk.bn3, relu=True, residual=idn)
                n_convs += 3
            elif kind == "BasicBlock":
                t, h1, w1, c1 = low.conv_bn(x, H, W, C, blk.conv1, blk.bn1, relu=True)
                x, H, W, C = low.conv_bn(t, h1, w1, c1, blk.conv2, blk.bn2, relu=True, residual=idn)
                n_convs += 2
            else:
                raise ValueError("b200 engine: unsupported residual block '{}'".format(kind))
    pooled = g.buffer("float16", False, C, rows_per_item=1)
    g.op(OP_AVGPOOL, [x, pooled, H * W, C])
    n_cls = model.fc.out_features
    logits = g.buffer("float32", False, n_cls, rows_per_item=1)
    g.linear(pooled, model.fc.weight.detach().double().numpy(), model.fc.bias.detach().double().numpy(),
             g.output(logits), out_f32=True)
    desc = dict(kind="graph", arch=type(model).__name__.lower(), convs=n_convs, num_classes=int(n_cls),
                input_dtype=input_dtype, output_dtype="float32", image_hw=list(image_hw), max_row_elems=0,
                implicit_conv=bool(implicit_conv))
    return PackedModel(native.MODEL_GRAPH, g.serialise(), desc)
