"""Model-repository ingestion for the b200 engine.

What the reference's Triton sidecar does with an endpoint's model file
(clearml_serving/engines/triton/triton_helper.py:91-194: fetch the registered model, look at its
`framework` tag, place it as model.pt / model.onnx / model.savedmodel / model.plan / model.bin and
write a config.pbtxt, :291-409) happens here IN PROCESS: the file is sniffed / matched against the
framework tag, lowered offline to a packed blob of this library's kernels (formats.py) and cached, so
an existing `clearml-serving model add --engine triton ...` registration can be served by
`engine_type="b200"` without re-registering the model.

Supported containers
  * TorchScript archives (`model.pt`, framework "pytorch"): torch.jit.load -> architecture recognised
    from the state_dict (torchvision ResNet family, HF BERT sequence classifiers) -> pack_resnet / pack_bert;
  * transformers `save_pretrained` folders;
  * XGBoost models saved as JSON or UBJSON (`Booster.save_model("m.json" | "m.ubj")`), the formats
    `XGBoostPreprocessRequest` loads at preprocess_service.py:475-476;
  * joblib / pickle sklearn estimators (`SKLearnPreprocessRequest`, preprocess_service.py:456-457);
  * a Triton model-repository folder `<name>/<version>/model.*` (the layout triton_helper.py:124-186 writes).
Refused loudly (no silent fallback): ONNX / TensorFlow / TensorRT containers, XGBoost's legacy binary format.
"""
import hashlib
import os
import re
import struct
import threading
import zipfile

import numpy as np

from . import formats

# triton_helper.py:159-186 (placement) + :378-385 (backend choice), restated as "which loader"
_FRAMEWORK_RULES = (
    (("pytorch", "torch", "caffe"), "torchscript"),
    (("xgboost",), "xgboost"),
    (("scikit", "sklearn", "joblib"), "sklearn"),
    (("onnx",), "onnx"),
    (("tensorflow", "keras"), "tensorflow"),
    (("tensorrt",), "tensorrt"),
)
_UNSUPPORTED = {
    "onnx": "ONNX graphs are not lowered by the b200 engine (export the torch module as TorchScript instead)",
    "tensorflow": "TensorFlow / Keras models are not supported by the b200 engine",
    "tensorrt": "TensorRT plans are device-specific binaries the b200 engine cannot read",
}


def loader_for_framework(framework):
    """framework tag of a registered model (`Model.framework`) -> loader name, or None (sniff the file)"""
    fw = str(framework or "").lower()
    for needles, loader in _FRAMEWORK_RULES:
        if any(n in fw for n in needles):
            return loader
    return None


# ------------------------------------------------------------------------------------------------
# UBJSON (https://ubjson.org, draft 12) -- the subset xgboost's `save_model("*.ubj")` emits: objects,
# arrays, strongly typed arrays (`[$<type>#<count>`), strings, all integer widths, float32/64, bool, null.
# ------------------------------------------------------------------------------------------------
_UBJ_NUM = {b"i": (">b", 1), b"U": (">B", 1), b"I": (">h", 2), b"l": (">i", 4), b"L": (">q", 8),
            b"d": (">f", 4), b"D": (">d", 8)}
_UBJ_NP = {b"i": ">i1", b"U": ">u1", b"I": ">i2", b"l": ">i4", b"L": ">i8", b"d": ">f4", b"D": ">f8"}


class _UbjReader(object):
    def __init__(self, data):
        self.b = memoryview(data)
        self.i = 0

    def _take(self, n):
        if self.i + n > len(self.b):
            raise ValueError("UBJSON: truncated input")
        v = self.b[self.i:self.i + n]
        self.i += n
        return v

    def _marker(self):
        return bytes(self._take(1))

    def _int(self):
        m = self._marker()
        if m not in (b"i", b"U", b"I", b"l", b"L"):
            raise ValueError("UBJSON: expected an integer length, got marker {!r}".format(m))
        fmt, n = _UBJ_NUM[m]
        return struct.unpack(fmt, self._take(n))[0]

    def _string(self):
        n = self._int()
        return bytes(self._take(n)).decode("utf-8")

    def value(self, marker=None):
        m = marker or self._marker()
        while m == b"N":   # no-op
            m = self._marker()
        if m in _UBJ_NUM:
            fmt, n = _UBJ_NUM[m]
            return struct.unpack(fmt, self._take(n))[0]
        if m == b"S":
            return self._string()
        if m == b"C":
            return bytes(self._take(1)).decode("latin-1")
        if m == b"T":
            return True
        if m == b"F":
            return False
        if m == b"Z":
            return None
        if m == b"H":   # high-precision number, kept as text
            return self._string()
        if m == b"[":
            return self._array()
        if m == b"{":
            return self._object()
        raise ValueError("UBJSON: unknown marker {!r} at byte {}".format(m, self.i - 1))

    def _container_header(self):
        typ, count = None, None
        if bytes(self.b[self.i:self.i + 1]) == b"$":
            self.i += 1
            typ = self._marker()
        if bytes(self.b[self.i:self.i + 1]) == b"#":
            self.i += 1
            count = self._int()
        elif typ is not None:
            raise ValueError("UBJSON: typed container without a count")
        return typ, count

    def _array(self):
        typ, count = self._container_header()
        if typ is not None and typ in _UBJ_NP:   # strongly typed numeric array: one bulk read
            dt = np.dtype(_UBJ_NP[typ])
            raw = self._take(count * dt.itemsize)
            return np.frombuffer(raw, dtype=dt).astype(dt.newbyteorder("=")).tolist()
        out = []
        if count is not None:
            for _ in range(count):
                out.append(self.value(typ))
            return out
        while True:
            m = self._marker()
            if m == b"]":
                return out
            out.append(self.value(m))

    def _object(self):
        typ, count = self._container_header()
        out = {}
        if count is not None:
            for _ in range(count):
                k = self._string()
                out[k] = self.value(typ)
            return out
        while True:
            if bytes(self.b[self.i:self.i + 1]) == b"}":
                self.i += 1
                return out
            k = self._string()
            out[k] = self.value()


def ubjson_loads(data):
    r = _UbjReader(data)
    v = r.value()
    return v


def ubjson_dumps(obj):
    """Writer used by the tests (and handy for converting a JSON model): emits the same subset,
    with strongly typed arrays for homogeneous int / float lists like xgboost does."""
    out = bytearray()

    def w_int(n):
        if -128 <= n <= 127:
            out.extend(b"i" + struct.pack(">b", n))
        elif 0 <= n <= 255:
            out.extend(b"U" + struct.pack(">B", n))
        elif -32768 <= n <= 32767:
            out.extend(b"I" + struct.pack(">h", n))
        elif -2 ** 31 <= n < 2 ** 31:
            out.extend(b"l" + struct.pack(">i", n))
        else:
            out.extend(b"L" + struct.pack(">q", n))

    def w_str(s):
        b = s.encode("utf-8")
        w_int(len(b))
        out.extend(b)

    def w(v):
        if v is None:
            out.extend(b"Z")
        elif v is True:
            out.extend(b"T")
        elif v is False:
            out.extend(b"F")
        elif isinstance(v, (int, np.integer)):
            w_int(int(v))
        elif isinstance(v, (float, np.floating)):
            out.extend(b"D" + struct.pack(">d", float(v)))
        elif isinstance(v, str):
            out.extend(b"S")
            w_str(v)
        elif isinstance(v, (list, tuple, np.ndarray)):
            seq = list(v)
            if seq and all(isinstance(x, (float, np.floating)) for x in seq):
                out.extend(b"[$d#")
                w_int(len(seq))
                out.extend(np.asarray(seq, dtype=">f4").tobytes())
            elif seq and all(isinstance(x, (int, np.integer)) and not isinstance(x, bool) for x in seq) and \
                    all(-2 ** 31 <= int(x) < 2 ** 31 for x in seq):
                out.extend(b"[$l#")
                w_int(len(seq))
                out.extend(np.asarray(seq, dtype=">i4").tobytes())
            else:
                out.extend(b"[")
                for x in seq:
                    w(x)
                out.extend(b"]")
        elif isinstance(v, dict):
            out.extend(b"{")
            for k, x in v.items():
                w_str(str(k))
                w(x)
            out.extend(b"}")
        else:
            raise TypeError("ubjson_dumps: unsupported type {}".format(type(v)))

    w(obj)
    return bytes(out)


# ------------------------------------------------------------------------------------------------
# TorchScript archives: recognise the architecture from the parameter names / shapes and rebuild the
# eager module the packers of formats.py understand (the weights ARE the model; no graph is traced).
# ------------------------------------------------------------------------------------------------
def _resnet_from_state_dict(sd):
    import torch
    import torchvision
    from torchvision.models.resnet import BasicBlock, Bottleneck
    layers = []
    for li in range(1, 5):
        idx = {int(m.group(1)) for k in sd for m in [re.match(r"layer{}\.(\d+)\.".format(li), k)] if m}
        if not idx:
            raise ValueError("b200 engine: TorchScript ResNet without layer{}".format(li))
        layers.append(max(idx) + 1)
    block = Bottleneck if "layer1.0.conv3.weight" in sd else BasicBlock
    width = int(sd["layer1.0.conv1.weight"].shape[0])
    kwargs = {}
    if block is Bottleneck and width != 64:   # wide / grouped variants keep the 64-base stem
        kwargs["width_per_group"] = width
    m = torchvision.models.resnet.ResNet(block, layers, num_classes=int(sd["fc.weight"].shape[0]), **kwargs)
    missing, unexpected = m.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()}, strict=False)
    bad = [k for k in list(missing) + list(unexpected) if not k.endswith("num_batches_tracked")]
    if bad:
        raise ValueError("b200 engine: TorchScript ResNet does not match torchvision's layout: {}".format(bad[:4]))
    return m.eval()


def _bert_from_state_dict(sd):
    import torch
    from transformers import BertConfig, BertForSequenceClassification
    p = "bert." if any(k.startswith("bert.") for k in sd) else ""
    emb = sd[p + "embeddings.word_embeddings.weight"]
    n_layers = 1 + max(int(m.group(1)) for k in sd for m in [re.search(r"encoder\.layer\.(\d+)\.", k)] if m)
    hidden = int(emb.shape[1])
    if hidden % 64:
        raise ValueError("b200 engine: BERT hidden size {} is not a multiple of the 64-wide heads".format(hidden))
    if "classifier.weight" not in sd:
        raise ValueError("b200 engine: TorchScript BERT without a `classifier` head is not supported")
    cfg = BertConfig(vocab_size=int(emb.shape[0]), hidden_size=hidden, num_hidden_layers=n_layers,
                     num_attention_heads=hidden // 64,
                     intermediate_size=int(sd[p + "encoder.layer.0.intermediate.dense.weight"].shape[0]),
                     max_position_embeddings=int(sd[p + "embeddings.position_embeddings.weight"].shape[0]),
                     type_vocab_size=int(sd[p + "embeddings.token_type_embeddings.weight"].shape[0]),
                     num_labels=int(sd["classifier.weight"].shape[0]))
    m = BertForSequenceClassification(cfg)
    missing, unexpected = m.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()}, strict=False)
    bad = [k for k in list(missing) + list(unexpected) if "position_ids" not in k and "token_type_ids" not in k]
    if bad:
        raise ValueError("b200 engine: TorchScript BERT does not match transformers' layout: {}".format(bad[:4]))
    return m.eval()


def lower_torchscript(path):
    """model.pt as the reference places it for Triton's libtorch backend (triton_helper.py:166-168)."""
    import torch
    try:
        sm = torch.jit.load(path, map_location="cpu")
    except Exception as ex:
        raise ValueError("b200 engine: '{}' is not a loadable TorchScript archive: {}".format(path, ex))
    sd = {k: v.detach() for k, v in sm.state_dict().items()}
    if "conv1.weight" in sd and "layer1.0.conv1.weight" in sd and "fc.weight" in sd:
        return formats.pack_resnet(_resnet_from_state_dict(sd))
    if any("encoder.layer.0.attention.self.query.weight" in k for k in sd):
        return formats.pack_bert(_bert_from_state_dict(sd))
    raise ValueError("b200 engine: TorchScript architecture not recognised (supported: torchvision ResNet family, "
                     "transformers BERT sequence classifiers); first parameters: {}".format(list(sd)[:4]))


# ------------------------------------------------------------------------------------------------
def _sniff(path):
    with open(path, "rb") as f:
        head = f.read(64)
    if head[:2] == b"PK":
        return "torchscript" if zipfile.is_zipfile(path) else "unknown"
    if head[:4] == b"binf" or head[:4] == b"bs64":
        return "xgboost-legacy"
    s = head.lstrip()
    if s[:1] == b"{":
        # JSON text continues with whitespace / a quote; UBJSON with a length marker (U/i/I/l/L), '$' or '#'
        nxt = s[1:2]
        return "xgboost-json" if nxt in (b'"', b" ", b"\n", b"\r", b"\t", b"}") else "xgboost-ubj"
    if head[:1] == b"\x80" or head[:2] == b"\x78\x9c" or head[:3] == b"ZF\x01":   # pickle / zlib-joblib
        return "sklearn"
    if head[:1] == b"\x08" and b"onnx" in head.lower():
        return "onnx"
    return "unknown"


def _find_in_repo_folder(path):
    """`<name>/<version>/model.*` or `<version>/model.*` or a folder holding one model file"""
    cands = []
    for root, _dirs, files in os.walk(path):
        for fn in files:
            if fn.startswith("model.") or fn.endswith((".pt", ".json", ".ubj", ".pkl", ".joblib")):
                cands.append(os.path.join(root, fn))
    if not cands:
        return None

    def version_key(p):   # Triton serves the highest numeric version folder
        parts = os.path.relpath(p, path).split(os.sep)
        nums = [int(x) for x in parts[:-1] if x.isdigit()]
        return (max(nums) if nums else -1, -len(parts))
    return sorted(cands, key=version_key)[-1]


def load_model(path, framework=None):
    """Local model file / folder (+ optional framework tag) -> formats.PackedModel."""
    path = str(path)
    if not os.path.exists(path):
        raise ValueError("b200 engine: model path '{}' does not exist".format(path))
    loader = loader_for_framework(framework)
    if loader in _UNSUPPORTED:
        raise ValueError("b200 engine: framework '{}': {}".format(framework, _UNSUPPORTED[loader]))
    if os.path.isdir(path):
        if os.path.exists(os.path.join(path, "config.json")):
            return formats.load_model_file(path)   # transformers folder
        inner = _find_in_repo_folder(path)
        if inner is None:
            raise ValueError("b200 engine: no model file under '{}'".format(path))
        return load_model(inner, framework)
    kind = _sniff(path)
    if kind == "xgboost-legacy":
        raise ValueError("b200 engine: XGBoost legacy binary models are not supported; re-save with "
                         "Booster.save_model('model.json') or 'model.ubj'")
    if kind == "onnx":
        raise ValueError("b200 engine: {}".format(_UNSUPPORTED["onnx"]))
    if loader == "torchscript" or (loader is None and kind == "torchscript"):
        return lower_torchscript(path)
    if kind == "xgboost-ubj" and loader in (None, "xgboost"):
        with open(path, "rb") as f:
            return formats.pack_xgboost_json(ubjson_loads(f.read()))
    if kind == "xgboost-json" and loader in (None, "xgboost"):
        return formats.pack_xgboost_json(path)
    if loader == "xgboost":
        raise ValueError("b200 engine: '{}' is neither an XGBoost JSON nor a UBJSON model".format(path))
    if kind in ("sklearn", "unknown") or loader == "sklearn":
        import joblib   # the reference's sklearn engine loads with joblib too (preprocess_service.py:456)
        try:
            obj = joblib.load(path)
        except Exception as ex:
            raise ValueError("b200 engine: cannot read model file '{}' ({}: {})".format(path, type(ex).__name__, ex))
        return formats.pack_sklearn(obj)
    raise ValueError("b200 engine: model file '{}' (framework {!r}) is not a supported container".format(path, framework))


# ------------------------------------------------------------------------------------------------
# auxiliary_cfg (config.pbtxt text or dict): validation + the keys this engine consumes
# ------------------------------------------------------------------------------------------------
def _pbtxt_top_level_keys(text):
    """names of the top-level fields of a config.pbtxt (enough for validation; no protobuf dependency)"""
    keys, depth, i, n = [], 0, 0, len(text)
    while i < n:
        c = text[i]
        if c == "#":
            while i < n and text[i] != "\n":
                i += 1
        elif c in "\"'":
            q = c
            i += 1
            while i < n and text[i] != q:
                i += 2 if text[i] == "\\" else 1
        elif c in "{[":
            depth += 1
        elif c in "}]":
            depth -= 1
        elif depth == 0 and (c.isalpha() or c == "_"):
            j = i
            while j < n and (text[j].isalnum() or text[j] == "_"):
                j += 1
            k = j
            while k < n and text[k] in " \t\r\n":
                k += 1
            if k < n and text[k] in ":{[":
                keys.append(text[i:j])
            i = j - 1
        i += 1
    return keys


def validate_auxiliary_cfg(aux):
    """model_request_processor.py:1464-1516 for this engine: the io description comes from the endpoint
    (`--input-name/--input-type/--input-size ...`), never from the pbtxt; `default_model_filename` is
    the system's to set.  Returns the list of top-level keys seen."""
    if not aux:
        return []
    if isinstance(aux, dict):
        keys = sorted({str(k).split(".")[0] for k in aux})
    else:
        keys = _pbtxt_top_level_keys(str(aux))
    if "input" in keys or "output" in keys:
        raise ValueError("b200 engine requires *manual* input/output specification, You input/output in your pbtxt, "
                         "please remove them and specify manually.")
    if "default_model_filename" in keys:
        raise ValueError("ERROR: You have `default_model_filename` in your config pbtxt, please remove it. "
                         "It will be added automatically by the system.")
    return keys


def parse_instance_group(aux):
    """`instance_group [{ count: 2  kind: KIND_GPU  gpus: [0, 1] }]` -> (count, gpus or None).  Triton runs `count`
    execution instances per listed GPU; here: `count` staging lanes (CUDA streams) per replica, replicas on `gpus`."""
    if not aux:
        return None, None
    if isinstance(aux, dict):
        ig = aux.get("instance_group")
        if isinstance(ig, (list, tuple)) and ig:
            ig = ig[0]
        if isinstance(ig, dict):
            gp = ig.get("gpus")
            return (int(ig["count"]) if ig.get("count") else None), ([int(g) for g in gp] if gp else None)
        cnt = aux.get("instance_group.count") or aux.get("instance_group.0.count")
        gp = aux.get("instance_group.gpus") or aux.get("instance_group.0.gpus")
        if isinstance(gp, str):
            gp = re.findall(r"\d+", gp)
        return (int(cnt) if cnt else None), ([int(g) for g in gp] if gp else None)
    m = re.search(r"instance_group\s*:?\s*\[\s*\{([^}]*)\}", str(aux), re.S)
    if not m:
        return None, None
    body = m.group(1)
    c = re.search(r"count\s*:\s*(\d+)", body)
    g = re.search(r"gpus\s*:\s*\[([^\]]*)\]", body)
    return (int(c.group(1)) if c else None), ([int(x) for x in re.findall(r"\d+", g.group(1))] if g else None)


# ------------------------------------------------------------------------------------------------
class ModelRepository(object):
    """Cache of packed models keyed by the content of the source file, refreshed when endpoints change:
    the in-process counterpart of `TritonHelper.model_service_update_step` (triton_helper.py:91-194)."""

    def __init__(self, resolver=None):
        self._resolver = resolver   # callable(model_id) -> (local_path, framework) or local_path
        self._lock = threading.Lock()
        self._packed = {}           # fingerprint -> PackedModel
        self._current = {}          # url -> (endpoint dict, fingerprint)

    @staticmethod
    def fingerprint(path):
        h = hashlib.sha256()
        if os.path.isdir(path):
            for root, _d, files in sorted(os.walk(path)):
                for fn in sorted(files):
                    st = os.stat(os.path.join(root, fn))
                    h.update("{}:{}:{}".format(os.path.relpath(os.path.join(root, fn), path), st.st_size,
                                               int(st.st_mtime)).encode())
        else:
            with open(path, "rb") as f:
                for blk in iter(lambda: f.read(1 << 20), b""):
                    h.update(blk)
        return h.hexdigest()

    def resolve(self, model_id):
        if self._resolver is None:
            return model_id, None
        r = self._resolver(model_id)
        return r if isinstance(r, tuple) else (r, None)

    def get(self, path, framework=None):
        fp = self.fingerprint(path)
        with self._lock:
            pm = self._packed.get(fp)
        if pm is None:
            pm = load_model(path, framework)
            with self._lock:
                self._packed[fp] = pm
        return pm, fp

    def update_step(self, active_endpoints, engine_types=("b200", "triton")):
        """active_endpoints: {url: ModelEndpoint-like (as_dict / dict)}.  Packs new / changed models, forgets
        removed ones; returns the urls whose engines must be rebuilt (empty list: nothing to do)."""
        changed = []
        seen = {}
        for url, ep in active_endpoints.items():
            d = ep.as_dict() if hasattr(ep, "as_dict") else dict(ep)
            if d.get("engine_type") not in engine_types or not d.get("model_id"):
                continue
            path, framework = self.resolve(d["model_id"])
            if not path or not os.path.exists(path):
                # like the reference: report and skip, the endpoint fails at first request instead
                print("Error retrieving model ID {} []".format(d["model_id"]))
                continue
            validate_auxiliary_cfg(d.get("auxiliary_cfg"))
            _pm, fp = self.get(path, framework)
            seen[url] = (d, fp)
            if self._current.get(url) != (d, fp):
                changed.append(url)
        removed = [u for u in self._current if u not in seen]
        live = {fp for _d, fp in seen.values()}
        with self._lock:
            for fp in [f for f in self._packed if f not in live]:
                del self._packed[fp]
        self._current = seen
        return changed + removed


_default_repo = None
_default_repo_lock = threading.Lock()


def default_repository():
    """process-wide packed-model cache: engines are rebuilt after every configuration change
    (model_request_processor.py:1026-1028), the lowering of an unchanged model file is not repeated"""
    global _default_repo
    with _default_repo_lock:
        if _default_repo is None:
            _default_repo = ModelRepository()
        return _default_repo
