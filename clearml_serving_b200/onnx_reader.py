"""Minimal reader of ONNX model files (the subset the engine lowers), without the `onnx` package.

The reference places `model.onnx` in the Triton repository folder for ONNX-Runtime (`framework` "onnx" ->
`platform: "onnxruntime_onnx"`, clearml_serving/engines/triton/triton_helper.py:169-171,378-385).  Here the file is parsed
directly -- ONNX is plain protobuf; the wire format needs ~100 lines -- into nodes / initializers / graph inputs, and
`model_repo.lower_onnx` lowers the graph onto this library's kernels.  Field numbers follow onnx/onnx.proto (IR version 3+):
ModelProto.graph = 7; GraphProto.node = 1, .initializer = 5, .input = 11, .output = 12; NodeProto.input = 1, .output = 2,
.op_type = 4, .attribute = 5; AttributeProto.name = 1, .f = 2, .i = 3, .s = 4, .t = 5, .floats = 7, .ints = 8;
TensorProto.dims = 1, .data_type = 2, .float_data = 4, .int32_data = 5, .int64_data = 7, .name = 8, .raw_data = 9.
"""
import struct

import numpy as np


class OnnxError(ValueError):
    pass


# TensorProto.DataType -> numpy
_DTYPES = {1: np.float32, 2: np.uint8, 3: np.int8, 5: np.int16, 6: np.int32, 7: np.int64, 9: np.bool_, 10: np.float16,
           11: np.float64, 12: np.uint32, 13: np.uint64}


def _varint(buf, pos):
    result, shift = 0, 0
    while True:
        if pos >= len(buf):
            raise OnnxError("onnx: truncated varint")
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 70:
            raise OnnxError("onnx: varint too long")


def _fields(buf):
    """yield (field number, wire type, value) of one message; length-delimited values are memoryviews"""
    buf = memoryview(buf)
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = bytes(buf[pos:pos + 8]); pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            if pos + ln > n:
                raise OnnxError("onnx: truncated field {}".format(field))
            v = buf[pos:pos + ln]; pos += ln
        elif wt == 5:
            v = bytes(buf[pos:pos + 4]); pos += 4
        else:
            raise OnnxError("onnx: unsupported wire type {}".format(wt))
        yield field, wt, v


def _signed(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _packed_varints(v):
    out, pos = [], 0
    while pos < len(v):
        x, pos = _varint(v, pos)
        out.append(_signed(x))
    return out


def _tensor(buf):
    dims, dtype, name, raw = [], 1, "", None
    floats, int32s, int64s = [], [], []
    for f, wt, v in _fields(buf):
        if f == 1:
            dims.extend(_packed_varints(v) if wt == 2 else [_signed(v)])
        elif f == 2:
            dtype = v
        elif f == 4:
            floats.append(np.frombuffer(v, "<f4") if wt == 2 else np.frombuffer(v, "<f4"))
        elif f == 5:
            int32s.extend(_packed_varints(v) if wt == 2 else [_signed(v)])
        elif f == 7:
            int64s.extend(_packed_varints(v) if wt == 2 else [_signed(v)])
        elif f == 8:
            name = bytes(v).decode()
        elif f == 9:
            raw = v
        elif f in (13, 14) and (wt != 0 or v != 0):
            raise OnnxError("onnx: tensor '{}' keeps its data in an external file (not supported)".format(name))
    if dtype not in _DTYPES:
        raise OnnxError("onnx: tensor '{}' has unsupported data type {}".format(name, dtype))
    dt = np.dtype(_DTYPES[dtype])
    if raw is not None:
        arr = np.frombuffer(raw, dt.newbyteorder("<"))
    elif floats:
        arr = np.concatenate(floats).astype(dt)
    elif int64s:
        arr = np.asarray(int64s, dt)
    elif int32s:
        arr = np.asarray(int32s, dt)   # also carries float16 / uint8 / bool payloads per the spec
    else:
        arr = np.zeros(0, dt)
    n = int(np.prod(dims)) if dims else arr.size
    if arr.size != n:
        raise OnnxError("onnx: tensor '{}' has {} elements for shape {}".format(name, arr.size, dims))
    return name, arr.reshape(dims)


def _attribute(buf):
    name, val = "", None
    ints, floats = [], []
    for f, wt, v in _fields(buf):
        if f == 1:
            name = bytes(v).decode()
        elif f == 2:
            val = struct.unpack("<f", v)[0]
        elif f == 3:
            val = _signed(v)
        elif f == 4:
            val = bytes(v)
        elif f == 5:
            val = _tensor(v)[1]
        elif f == 7:
            floats.extend(np.frombuffer(v, "<f4").tolist() if wt == 2 else [struct.unpack("<f", v)[0]])
        elif f == 8:
            ints.extend(_packed_varints(v) if wt == 2 else [_signed(v)])
    if ints:
        val = ints
    elif floats:
        val = floats
    return name, val


class Node(object):
    __slots__ = ("op_type", "inputs", "outputs", "attrs", "name")

    def __init__(self):
        self.op_type, self.inputs, self.outputs, self.attrs, self.name = "", [], [], {}, ""

    def __repr__(self):
        return "{}({} -> {})".format(self.op_type, ",".join(self.inputs), ",".join(self.outputs))


def _node(buf):
    n = Node()
    for f, wt, v in _fields(buf):
        if f == 1:
            n.inputs.append(bytes(v).decode())
        elif f == 2:
            n.outputs.append(bytes(v).decode())
        elif f == 3:
            n.name = bytes(v).decode()
        elif f == 4:
            n.op_type = bytes(v).decode()
        elif f == 5:
            k, a = _attribute(v)
            n.attrs[k] = a
    return n


def _value_info(buf):
    """-> (name, numpy dtype or None, shape with None for symbolic dims)"""
    name, dtype, shape = "", None, None
    for f, wt, v in _fields(buf):
        if f == 1:
            name = bytes(v).decode()
        elif f == 2:
            for f2, _, v2 in _fields(v):               # TypeProto
                if f2 != 1:
                    continue
                for f3, wt3, v3 in _fields(v2):        # TypeProto.Tensor
                    if f3 == 1:
                        dtype = _DTYPES.get(v3)
                    elif f3 == 2:
                        shape = []
                        for f4, _, v4 in _fields(v3):  # TensorShapeProto.dim
                            dim = None
                            for f5, wt5, v5 in _fields(v4):
                                if f5 == 1:
                                    dim = _signed(v5)
                            shape.append(dim)
    return name, dtype, shape


class Graph(object):
    def __init__(self):
        self.nodes, self.initializers, self.inputs, self.outputs = [], {}, [], []


def load(data):
    """bytes of a .onnx file -> Graph (nodes in file order = topological order, initializers by name, graph inputs that are
    not initializers)"""
    if isinstance(data, str):
        with open(data, "rb") as f:
            data = f.read()
    graph_buf = None
    for f, wt, v in _fields(data):
        if f == 7 and wt == 2:
            graph_buf = v
    if graph_buf is None:
        raise OnnxError("onnx: no graph in the model file")
    g = Graph()
    for f, wt, v in _fields(graph_buf):
        if f == 1:
            g.nodes.append(_node(v))
        elif f == 5:
            name, arr = _tensor(v)
            g.initializers[name] = arr
        elif f == 11:
            g.inputs.append(_value_info(v))
        elif f == 12:
            g.outputs.append(_value_info(v))
    g.inputs = [i for i in g.inputs if i[0] not in g.initializers]
    if not g.nodes:
        raise OnnxError("onnx: empty graph")
    return g
