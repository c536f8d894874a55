"""Binary tensor frames on the REST edge.

The reference's route takes `Union[bytes, Dict]` bodies (clearml_serving/serving/main.py:197) and its PyTorch
example already ships raw image bytes (examples/pytorch/preprocess.py:19-25); JSON-encoding 150 528 floats per
ResNet image is what caps REST-level throughput.  A frame carries the request tensors of ONE
request exactly as `TritonPreprocessRequest.process` would marshal them (preprocess_service.py:385-406: one array
per declared input, batch dimension first) with no per-element boxing: the engine's collate step reads the
payload in place (`np.frombuffer`, zero copy) and the reply goes back in the same framing.

Frame (little endian), Content-Type `application/x-b200-tensors`:
    0   4s  magic  "B2ST"
    4   u8  version (1)
    5   u8  n_tensors (1..8)
    6   u16 reserved (0)
    8   n_tensors x { u8 dtype, u8 ndim (0..8), u16 reserved, u32 shape[ndim] }
    ..  zero padding to a multiple of 8
    ..  tensor payloads, C order, each padded to a multiple of 8 bytes
dtype codes are the library's (include/b200serve.h `b2s_dtype`): the universe the reference's Triton client can
put on the wire (preprocess_service.py:271-282) plus float16.
"""
import struct

import numpy as np

MEDIA_TYPE = "application/x-b200-tensors"
MAGIC = b"B2ST"
VERSION = 1
MAX_TENSORS = 8
MAX_NDIM = 8

# b2s_dtype codes (include/b200serve.h)
_CODE_TO_DTYPE = {0: np.float32, 1: np.float64, 2: np.int32, 3: np.int64, 4: np.uint8, 5: np.int8, 6: np.bool_,
                  7: np.uint64, 8: np.float16, 9: np.uint32}
_DTYPE_TO_CODE = {np.dtype(v): k for k, v in _CODE_TO_DTYPE.items()}


class WireError(ValueError):
    pass


def is_tensor_frame(body):
    return isinstance(body, (bytes, bytearray, memoryview)) and len(body) >= 8 and bytes(body[:4]) == MAGIC


def encode_tensors(tensors):
    """list of array-likes -> one frame (bytes)"""
    arrays = []
    for t in tensors:
        a = np.asarray(t)
        if a.dtype.byteorder == ">":
            a = a.astype(a.dtype.newbyteorder("<"))
        if a.dtype not in _DTYPE_TO_CODE:
            raise WireError("tensor frame: dtype {} is not supported".format(a.dtype))
        if a.ndim > MAX_NDIM:
            raise WireError("tensor frame: {} dimensions (max {})".format(a.ndim, MAX_NDIM))
        arrays.append(a if a.flags.c_contiguous else np.ascontiguousarray(a))   # (0-d stays 0-d)
    if not 1 <= len(arrays) <= MAX_TENSORS:
        raise WireError("tensor frame: {} tensors (1..{})".format(len(arrays), MAX_TENSORS))
    head = bytearray(struct.pack("<4sBBH", MAGIC, VERSION, len(arrays), 0))
    for a in arrays:
        head += struct.pack("<BBH", _DTYPE_TO_CODE[a.dtype], a.ndim, 0)
        head += struct.pack("<{}I".format(a.ndim), *a.shape)
    head += b"\0" * (-len(head) % 8)
    parts = [bytes(head)]
    for a in arrays:
        raw = a.tobytes()
        parts.append(raw)
        if len(raw) % 8:
            parts.append(b"\0" * (-len(raw) % 8))
    return b"".join(parts)


def decode_tensors(body):
    """frame -> list of read-only ndarrays that VIEW `body` (no copy); raises WireError on any malformed field"""
    mv = memoryview(body)
    if len(mv) < 8 or bytes(mv[:4]) != MAGIC:
        raise WireError("tensor frame: bad magic")
    _m, version, n, _r = struct.unpack_from("<4sBBH", mv, 0)
    if version != VERSION:
        raise WireError("tensor frame: version {} is not supported".format(version))
    if not 1 <= n <= MAX_TENSORS:
        raise WireError("tensor frame: {} tensors (1..{})".format(n, MAX_TENSORS))
    off = 8
    specs = []
    for _ in range(n):
        if off + 4 > len(mv):
            raise WireError("tensor frame: truncated header")
        code, ndim, _r = struct.unpack_from("<BBH", mv, off)
        off += 4
        if code not in _CODE_TO_DTYPE or ndim > MAX_NDIM:
            raise WireError("tensor frame: bad dtype code {} / ndim {}".format(code, ndim))
        if off + 4 * ndim > len(mv):
            raise WireError("tensor frame: truncated header")
        shape = struct.unpack_from("<{}I".format(ndim), mv, off)
        off += 4 * ndim
        specs.append((np.dtype(_CODE_TO_DTYPE[code]), shape))
    off += -off % 8
    out = []
    for dt, shape in specs:
        count = 1
        for s in shape:
            count *= int(s)
        nbytes = count * dt.itemsize
        if off + nbytes > len(mv):
            raise WireError("tensor frame: payload truncated ({} bytes missing)".format(off + nbytes - len(mv)))
        out.append(np.frombuffer(mv, dtype=dt, count=count, offset=off).reshape(shape))
        off += nbytes + (-nbytes % 8)
    return out
