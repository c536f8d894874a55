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
  * LightGBM text models (`Booster.save_model("m.txt")`, `LightGBMPreprocessRequest`, preprocess_service.py:486-501);
  * joblib / pickle sklearn estimators (`SKLearnPreprocessRequest`, preprocess_service.py:456-457);
  * a Triton model-repository folder `<name>/<version>/model.*` (the layout triton_helper.py:124-186 writes).
  * ONNX files of residual convolutional networks (torchvision ResNet family as `torch.onnx.export` writes them:
    Conv with the BatchNorm already folded, Relu, MaxPool, Add, GlobalAveragePool, Flatten, Gemm), read without the
    `onnx` package (onnx_reader.py) -- the file the reference places as model.onnx for Triton's ONNX-Runtime backend
    (triton_helper.py:169-171).
Refused loudly (no silent fallback): other ONNX graphs, TensorFlow / TensorRT containers, XGBoost's legacy binary format.
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
    (("lightgbm",), "lightgbm"),
    (("scikit", "sklearn", "joblib"), "sklearn"),
    (("onnx",), "onnx"),
    (("tensorflow", "keras"), "tensorflow"),
    (("tensorrt",), "tensorrt"),
)
_UNSUPPORTED = {
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
# ONNX: residual convolutional networks.  The graph is walked in file (= topological) order and matched against the
# torchvision block structure -- stem Conv 7x7/2 (+Relu) + MaxPool 3x3/2, then blocks `Conv Relu Conv [Relu Conv]
# [downsample Conv] Add Relu`, then GlobalAveragePool, Flatten, Gemm -- and rebuilt as the eager module formats.pack_resnet
# lowers (the exporter has folded every BatchNorm into its convolution: the rebuilt BatchNorms are identities carrying
# the convolution bias), so an ONNX file and the TorchScript / eager form of the same network give the same op list.
# ------------------------------------------------------------------------------------------------
def _onnx_conv_spec(g, node):
    a = node.attrs
    w = g.initializers.get(node.inputs[1]) if len(node.inputs) > 1 else None
    if w is None or w.ndim != 4:
        raise ValueError("b200 engine: ONNX Conv '{}' without a constant 4-D weight".format(node.name or node.outputs[0]))
    b = g.initializers.get(node.inputs[2]) if len(node.inputs) > 2 else None
    ks = list(a.get("kernel_shape", w.shape[2:]))
    st = list(a.get("strides", [1, 1]))
    pads = list(a.get("pads", [0, 0, 0, 0]))
    if a.get("group", 1) != 1 or any(d != 1 for d in a.get("dilations", [1, 1])) or ks[0] != ks[1] or st[0] != st[1] or \
            len(set(pads)) != 1 or a.get("auto_pad", b"NOTSET") not in (b"NOTSET", "NOTSET"):
        raise ValueError("b200 engine: ONNX Conv '{}': grouped / dilated / asymmetric convolutions are not supported".format(
            node.name or node.outputs[0]))
    return dict(w=np.array(w, np.float32), b=None if b is None else np.array(b, np.float32), k=int(ks[0]),
                stride=int(st[0]), pad=int(pads[0]))


def lower_onnx(data):
    """bytes / path of an ONNX file -> formats.PackedModel (torchvision-style ResNet: Bottleneck or BasicBlock stages)"""
    import torch
    import torchvision
    from torchvision.models.resnet import BasicBlock, Bottleneck
    from . import onnx_reader
    try:
        g = onnx_reader.load(data)
    except onnx_reader.OnnxError as ex:
        raise ValueError("b200 engine: {}".format(ex))
    if len(g.inputs) != 1 or len(g.outputs) != 1:
        raise ValueError("b200 engine: ONNX graphs with {} inputs / {} outputs are not supported".format(len(g.inputs), len(g.outputs)))
    in_name, in_dtype, in_shape = g.inputs[0]
    if in_dtype not in (np.float32, np.uint8) or not in_shape or len(in_shape) != 4 or in_shape[1] not in (1, 3, 4):
        raise ValueError("b200 engine: ONNX input must be NCHW float32 / uint8 images, got {} {}".format(in_dtype, in_shape))
    consumers = {}
    for n in g.nodes:
        for t in n.inputs:
            consumers.setdefault(t, []).append(n)
    producer = {t: n for n in g.nodes for t in n.outputs}

    def only(t, op=None):
        c = consumers.get(t, [])
        if len(c) != 1 or (op and c[0].op_type != op):
            raise ValueError("b200 engine: ONNX graph is not a residual conv net (tensor '{}' feeds {})".format(
                t, [x.op_type for x in c]))
        return c[0]

    unsupported = sorted({n.op_type for n in g.nodes} - {"Conv", "Relu", "MaxPool", "Add", "GlobalAveragePool", "Flatten", "Gemm"})
    if unsupported:
        raise ValueError("b200 engine: ONNX operators {} are not lowered (supported graphs: residual convolutional networks; "
                         "export transformer encoders as TorchScript)".format(unsupported))
    # ---- stem
    n = only(in_name, "Conv")
    stem = _onnx_conv_spec(g, n)
    n = only(n.outputs[0], "Relu")
    n = only(n.outputs[0], "MaxPool")
    a = n.attrs
    if list(a.get("kernel_shape", [])) != [3, 3] or list(a.get("strides", [])) != [2, 2] or list(a.get("pads", [])) != [1, 1, 1, 1] or a.get("ceil_mode", 0):
        raise ValueError("b200 engine: ONNX MaxPool must be 3x3 stride 2 pad 1")
    x = n.outputs[0]
    # ---- residual blocks
    blocks = []
    while True:
        users = consumers.get(x, [])
        if len(users) == 1 and users[0].op_type == "GlobalAveragePool":
            break
        convs = [u for u in users if u.op_type == "Conv"]
        adds = [u for u in users if u.op_type == "Add"]
        if not convs or len(convs) + len(adds) != len(users) or len(convs) > 2 or len(adds) > 1 or len(convs) + len(adds) != 2:
            raise ValueError("b200 engine: ONNX graph is not a residual conv net (block input '{}' feeds {})".format(
                x, [u.op_type for u in users]))
        main = [_onnx_conv_spec(g, convs[0])]   # file order: the main path's first convolution precedes the downsample one
        t = convs[0].outputs[0]
        while True:
            nx = only(t)
            if nx.op_type == "Relu":
                nx2 = only(nx.outputs[0], "Conv")
                main.append(_onnx_conv_spec(g, nx2))
                t = nx2.outputs[0]
            elif nx.op_type == "Add":
                add = nx
                break
            else:
                raise ValueError("b200 engine: ONNX graph is not a residual conv net ('{}' after a convolution)".format(nx.op_type))
        other = [i for i in add.inputs if i != t]
        if len(other) != 1:
            raise ValueError("b200 engine: ONNX Add with unexpected operands")
        down = None
        if other[0] != x:
            dn = producer.get(other[0])
            if dn is None or dn.op_type != "Conv" or dn.inputs[0] != x or len(convs) != 2 or dn is not convs[1]:
                raise ValueError("b200 engine: ONNX residual branch is neither the block input nor one convolution of it")
            down = _onnx_conv_spec(g, dn)
        elif len(convs) != 1:
            raise ValueError("b200 engine: ONNX block input feeds two convolutions but the identity is not a convolution")
        blocks.append(dict(main=main, down=down))
        x = only(add.outputs[0], "Relu").outputs[0]
        if len(blocks) > 512:
            raise ValueError("b200 engine: ONNX graph too deep")
    n = only(x, "GlobalAveragePool")
    n = only(n.outputs[0], "Flatten")
    n = only(n.outputs[0], "Gemm")
    a = n.attrs
    fw, fb = g.initializers.get(n.inputs[1]), g.initializers.get(n.inputs[2]) if len(n.inputs) > 2 else None
    if fw is None or a.get("alpha", 1.0) != 1.0 or a.get("beta", 1.0) != 1.0 or a.get("transA", 0):
        raise ValueError("b200 engine: ONNX Gemm must be a plain fully connected layer with constant weights")
    fw = np.array(fw, np.float32) if a.get("transB", 0) else np.array(fw, np.float32).T
    if n.outputs[0] != g.outputs[0][0] or not blocks:
        raise ValueError("b200 engine: ONNX graph does not end in the classifier")
    # ---- rebuild the torchvision module
    n_main = {len(b["main"]) for b in blocks}
    if n_main not in ({2}, {3}):
        raise ValueError("b200 engine: ONNX residual blocks must all have 2 (BasicBlock) or 3 (Bottleneck) convolutions")
    block = Bottleneck if n_main == {3} else BasicBlock
    layers, cur = [], 0
    for i, b in enumerate(blocks):
        if b["down"] is not None and i > 0:
            layers.append(cur)
            cur = 0
        cur += 1
    layers.append(cur)
    if len(layers) != 4:
        raise ValueError("b200 engine: ONNX conv net has {} stages; torchvision-style ResNets have 4".format(len(layers)))
    kwargs = {}
    width = int(blocks[0]["main"][0]["w"].shape[0])
    if block is Bottleneck and width != 64:
        kwargs["width_per_group"] = width
    m = torchvision.models.resnet.ResNet(block, layers, num_classes=int(fw.shape[0]), **kwargs)

    def put(conv, bn, spec, what):
        if tuple(conv.weight.shape) != tuple(spec["w"].shape) or conv.stride[0] != spec["stride"] or conv.padding[0] != spec["pad"]:
            raise ValueError("b200 engine: ONNX {} {} stride {} pad {} does not match torchvision's {} stride {} pad {}".format(
                what, spec["w"].shape, spec["stride"], spec["pad"], tuple(conv.weight.shape), conv.stride[0], conv.padding[0]))
        with torch.no_grad():
            conv.weight.copy_(torch.from_numpy(spec["w"]))
            bn.weight.fill_(1.0)
            bn.running_mean.zero_()
            bn.running_var.fill_(1.0)
            bn.bias.copy_(torch.from_numpy(spec["b"]) if spec["b"] is not None else torch.zeros_like(bn.bias))
        bn.eps = 0.0   # identity: y = x + bias, exactly
    put(m.conv1, m.bn1, stem, "stem")
    mods = [blk for layer in (m.layer1, m.layer2, m.layer3, m.layer4) for blk in layer]
    for i, (blk, b) in enumerate(zip(mods, blocks)):
        names = ("conv1", "conv2", "conv3")[:len(b["main"])]
        for nm, spec in zip(names, b["main"]):
            put(getattr(blk, nm), getattr(blk, nm.replace("conv", "bn")), spec, "block {} {}".format(i, nm))
        if (blk.downsample is None) != (b["down"] is None):
            raise ValueError("b200 engine: ONNX block {} downsample branch does not match torchvision's layout".format(i))
        if b["down"] is not None:
            put(blk.downsample[0], blk.downsample[1], b["down"], "block {} downsample".format(i))
    with torch.no_grad():
        m.fc.weight.copy_(torch.from_numpy(np.ascontiguousarray(fw)))
        m.fc.bias.copy_(torch.from_numpy(np.array(fb, np.float32)) if fb is not None else torch.zeros_like(m.fc.bias))
    hw = (int(in_shape[2]), int(in_shape[3])) if in_shape[2] and in_shape[3] else (224, 224)
    return formats.pack_resnet(m.eval(), input_dtype="uint8" if in_dtype == np.uint8 else "float32", image_hw=hw)


# ------------------------------------------------------------------------------------------------
def _sniff(path):
    with open(path, "rb") as f:
        head = f.read(256)
    if head[:2] == b"PK":
        return "torchscript" if zipfile.is_zipfile(path) else "unknown"
    if head[:4] == b"bs64" or formats.looks_like_xgboost_legacy_binary(head):
        return "xgboost-legacy"
    s = head.lstrip()
    if s[:5] == b"tree\n" or s[:6] == b"tree\r\n":
        return "lightgbm"   # Booster.save_model text: "tree", "version=v3", ...
    if s[:1] == b"{":
        # JSON text continues with whitespace / a quote; UBJSON with a length marker (U/i/I/l/L), '$' or '#'
        nxt = s[1:2]
        return "xgboost-json" if nxt in (b'"', b" ", b"\n", b"\r", b"\t", b"}") else "xgboost-ubj"
    if head[:1] == b"\x80" or head[:2] == b"\x78\x9c" or head[:3] == b"ZF\x01":   # pickle / zlib-joblib
        return "sklearn"
    if head[:1] == b"\x08" and (path.endswith(".onnx") or b"onnx" in head.lower() or b"pytorch" in head.lower()):
        return "onnx"   # ModelProto: field 1 (ir_version) first, then producer_name
    return "unknown"


def _find_in_repo_folder(path):
    """`<name>/<version>/model.*` or `<version>/model.*` or a folder holding one model file"""
    cands = []
    for root, _dirs, files in os.walk(path):
        for fn in files:
            if fn.startswith("model.") or fn.endswith((".pt", ".onnx", ".json", ".ubj", ".txt", ".pkl", ".joblib")):
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
    if kind == "xgboost-legacy" and loader in (None, "xgboost"):
        return formats.pack_xgboost_legacy_binary(path)   # what examples/xgboost/train_model.py:28 writes
    if loader == "onnx" or (loader is None and kind == "onnx"):
        return lower_onnx(path)
    if loader == "torchscript" or (loader is None and kind == "torchscript"):
        return lower_torchscript(path)
    if kind == "xgboost-ubj" and loader in (None, "xgboost"):
        with open(path, "rb") as f:
            return formats.pack_xgboost_json(ubjson_loads(f.read()))
    if kind == "xgboost-json" and loader in (None, "xgboost"):
        return formats.pack_xgboost_json(path)
    if loader == "xgboost":
        raise ValueError("b200 engine: '{}' is neither an XGBoost JSON nor a UBJSON model".format(path))
    if loader == "lightgbm" or (loader is None and kind == "lightgbm"):
        return formats.pack_lightgbm_text(path)
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
        resolver = self._resolver
        if resolver is None:   # the engines' own resolver (a model registry lookup installed by the host application)
            from .preprocess_service import BasePreprocessRequest
            resolver = BasePreprocessRequest._model_resolver
        if resolver is None:
            return model_id, None
        r = resolver(model_id)
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
