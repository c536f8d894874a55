"""Pure-numpy interpreter of the packed forest blob (csrc/forest.cu layout). TEST INFRASTRUCTURE:
lets the CPU suite check formats.pack_forest without a GPU by executing exactly the rules the
kernel executes (split `x < thr` in fp32, NaN -> default_left bit, children adjacent, sequential
accumulation in tree order in fp32 or fp64)."""
import struct

import numpy as np


def decode(blob):
    (magic, version, n_trees, n_features, n_nodes, feat_bits, acc_mode, n_leaf64, link, _r1, base,
     divisor) = struct.unpack_from("<4sIIIIIIIIIdd", blob, 0)
    assert magic == b"B2SF" and version == 1
    o = 56
    toff = np.frombuffer(blob, dtype="<u4", count=n_trees + 1, offset=o)
    o += ((n_trees + 1) * 4 + 7) // 8 * 8
    nodes = np.frombuffer(blob, dtype="<u8", count=n_nodes, offset=o)
    o += n_nodes * 8
    leaf64 = np.frombuffer(blob, dtype="<f8", count=n_leaf64, offset=o)
    assert o + n_leaf64 * 8 == len(blob)
    return dict(n_trees=n_trees, n_features=n_features, feat_bits=feat_bits, acc_mode=acc_mode, link=link,
                base=base, divisor=divisor, toff=toff, val=(nodes & np.uint64(0xffffffff)).astype(np.uint32),
                meta=(nodes >> np.uint64(32)).astype(np.uint32), leaf64=leaf64)


def predict(blob, X):
    d = decode(blob)
    X = np.ascontiguousarray(X, dtype=np.float32)
    fb = d["feat_bits"]
    fmask = (1 << fb) - 1
    f64 = d["acc_mode"] == 1
    out = np.empty(X.shape[0], dtype=np.float64 if f64 else np.float32)
    val_f32 = d["val"].view(np.float32)
    for i in range(X.shape[0]):
        acc = np.float64(d["base"]) if f64 else np.float32(d["base"])
        for t in range(d["n_trees"]):
            b = int(d["toff"][t])
            nid = 0
            while True:
                meta = int(d["meta"][b + nid])
                left = meta >> (fb + 1)
                if left == 0:
                    break
                x = X[i, meta & fmask]
                go_left = bool((meta >> fb) & 1) if np.isnan(x) else bool(x < val_f32[b + nid])
                nid = left + (0 if go_left else 1)
            if f64:
                acc = np.float64(acc + d["leaf64"][int(d["val"][b + nid])])
            else:
                acc = np.float32(acc + val_f32[b + nid])
        out[i] = (acc / np.float64(d["divisor"])) if f64 else acc
    if d["link"] == 1:   # xgboost's fp32 logistic transform (the kernel's forest_link_f32)
        from oracle import oracle as orc
        out = orc.xgb_sigmoid(out)
    return out
