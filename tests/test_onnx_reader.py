"""onnx_reader.py: the protobuf wire-format subset, fed hand-assembled messages (field numbers from onnx/onnx.proto) so that
every storage form a TensorProto may use is covered -- raw_data, float_data, int64_data, packed and unpacked repeated
fields -- plus the refusals (external data, unknown dtype, truncation)."""
import struct

import numpy as np
import pytest

from clearml_serving_b200 import onnx_reader as R


def _varint(v):
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _key(field, wt):
    return _varint((field << 3) | wt)


def _ld(field, payload):            # length-delimited
    return _key(field, 2) + _varint(len(payload)) + payload


def _vi(field, v):                  # varint
    return _key(field, 0) + _varint(v)


def _tensor(name, dims, dtype, raw=None, floats=None, int64s=None, packed_dims=True, extra=b""):
    msg = b""
    if packed_dims:
        msg += _ld(1, b"".join(_varint(d) for d in dims))
    else:
        msg += b"".join(_vi(1, d) for d in dims)
    msg += _vi(2, dtype) + _ld(8, name.encode())
    if raw is not None:
        msg += _ld(9, raw)
    if floats is not None:
        msg += _ld(4, np.asarray(floats, "<f4").tobytes())
    if int64s is not None:
        msg += b"".join(_vi(7, v) for v in int64s)          # unpacked repeated int64
    return msg + extra


def _attr_ints(name, ints, packed=True):
    body = _ld(1, name.encode())
    body += _ld(8, b"".join(_varint(i) for i in ints)) if packed else b"".join(_vi(8, i) for i in ints)
    return body


def _node(op, inputs, outputs, attrs=()):
    msg = b"".join(_ld(1, i.encode()) for i in inputs) + b"".join(_ld(2, o.encode()) for o in outputs) + _ld(4, op.encode())
    return msg + b"".join(_ld(5, a) for a in attrs)


def _value_info(name, elem_type, dims):
    shape = b"".join(_ld(1, _vi(1, d) if isinstance(d, int) else _ld(2, d.encode())) for d in dims)
    return _ld(1, name.encode()) + _ld(2, _ld(1, _vi(1, elem_type) + _ld(2, shape)))


def _model(nodes, inits, inputs, outputs):
    graph = b"".join(_ld(1, n) for n in nodes) + b"".join(_ld(5, t) for t in inits) + \
        b"".join(_ld(11, i) for i in inputs) + b"".join(_ld(12, o) for o in outputs)
    return _vi(1, 8) + _ld(2, b"pytorch") + _ld(7, graph)


def test_every_tensor_storage_form_and_attribute_kind():
    w = np.arange(24, dtype=np.float32).reshape(2, 3, 2, 2)
    inits = [_tensor("w_raw", w.shape, 1, raw=w.tobytes()),
             _tensor("w_float_data", (2, 2), 1, floats=[1.5, -2.0, 0.25, 8.0], packed_dims=False),
             _tensor("idx", (3,), 7, int64s=[5, -1, 1 << 40]),
             _tensor("half", (2,), 10, raw=np.array([0.5, -3.0], np.float16).tobytes()),
             _tensor("scalar", (), 1, raw=struct.pack("<f", 2.5))]
    f_attr = _ld(1, b"alpha") + _key(2, 5) + struct.pack("<f", 0.75)
    i_attr = _ld(1, b"axis") + _vi(3, -1)
    s_attr = _ld(1, b"auto_pad") + _ld(4, b"NOTSET")
    t_attr = _ld(1, b"value") + _ld(5, _tensor("", (2,), 6, raw=np.array([7, 9], np.int32).tobytes()))
    nodes = [_node("Conv", ["x", "w_raw"], ["y"], [_attr_ints("pads", [1, 1, 1, 1]), _attr_ints("strides", [2, 2], packed=False),
                                                   f_attr, i_attr, s_attr, t_attr]),
             _node("Relu", ["y"], ["z"])]
    g = R.load(_model(nodes, inits, [_value_info("x", 1, ["batch", 3, 8, 8]), _value_info("w_raw", 1, [2, 3, 2, 2])],
                      [_value_info("z", 1, ["batch", 2, 4, 4])]))
    assert [n.op_type for n in g.nodes] == ["Conv", "Relu"] and g.nodes[0].inputs == ["x", "w_raw"] and g.nodes[1].outputs == ["z"]
    a = g.nodes[0].attrs
    assert a["pads"] == [1, 1, 1, 1] and a["strides"] == [2, 2] and a["alpha"] == 0.75 and a["axis"] == -1 and a["auto_pad"] == b"NOTSET"
    assert a["value"].dtype == np.int32 and a["value"].tolist() == [7, 9]
    assert np.array_equal(g.initializers["w_raw"], w)
    assert np.array_equal(g.initializers["w_float_data"], [[1.5, -2.0], [0.25, 8.0]])
    assert g.initializers["idx"].dtype == np.int64 and g.initializers["idx"].tolist() == [5, -1, 1 << 40]
    assert g.initializers["half"].dtype == np.float16 and g.initializers["half"].tolist() == [0.5, -3.0]
    assert g.initializers["scalar"].shape == () and float(g.initializers["scalar"]) == 2.5
    assert g.inputs == [("x", np.float32, [None, 3, 8, 8])]        # initializers listed as inputs (IR < 4 style) are dropped
    assert g.outputs == [("z", np.float32, [None, 2, 4, 4])]


@pytest.mark.parametrize("build,msg", [
    (lambda: _model([_node("Relu", ["x"], ["y"])], [_tensor("w", (2,), 1, raw=b"\0" * 4)], [], []), "elements for shape"),
    (lambda: _model([_node("Relu", ["x"], ["y"])], [_tensor("w", (1,), 16, raw=b"\0\0")], [], []), "unsupported data type"),
    (lambda: _model([_node("Relu", ["x"], ["y"])], [_tensor("w", (1,), 1, extra=_vi(14, 1))], [], []), "external"),
    (lambda: _model([], [], [], []), "empty graph"),
    (lambda: _vi(1, 8) + _ld(2, b"pytorch"), "no graph"),
    (lambda: _model([_node("Relu", ["x"], ["y"])], [], [], [])[:-3], "truncated"),
])
def test_malformed_or_unsupported_files_raise_onnx_errors(build, msg):
    with pytest.raises(R.OnnxError, match=msg):
        R.load(build())
