"""Binary tensor frames on the REST edge (reference body contract main.py:197, gzip :32-50)."""
import gzip

import numpy as np
import pytest

from clearml_serving_b200 import ModelEndpoint, wire
from clearml_serving_b200.model_request_processor import ModelRequestProcessor
from tests.fakes import FakeModel, make_fake_engine


def test_frame_round_trip_is_zero_copy_and_exact():
    rng = np.random.default_rng(0)
    tensors = [rng.standard_normal((2, 3, 5)).astype(np.float32), rng.integers(0, 30522, (1, 17)).astype(np.int32),
               np.arange(7, dtype=np.uint8), np.array(3.5, dtype=np.float64), rng.standard_normal((4, 1)).astype(np.float16),
               np.array([[True, False]]), np.array([2 ** 63], dtype=np.uint64), np.zeros((0, 4), np.int64)]
    frame = wire.encode_tensors(tensors)
    assert wire.is_tensor_frame(frame) and len(frame) % 8 == 0
    back = wire.decode_tensors(frame)
    assert len(back) == len(tensors)
    for a, b in zip(tensors, back):
        assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b)
        assert not b.flags.owndata and not b.flags.writeable     # views into the request body


@pytest.mark.parametrize("mutate", [
    lambda f: b"XXXX" + f[4:],                      # magic
    lambda f: f[:4] + b"\x02" + f[5:],              # version
    lambda f: f[:5] + b"\x00" + f[6:],              # zero tensors
    lambda f: f[:8] + b"\x63" + f[9:],              # dtype code
    lambda f: f[:9] + b"\x09" + f[10:],             # ndim
    lambda f: f[:-8],                               # payload truncated
    lambda f: f[:10],                               # header truncated
])
def test_malformed_frames_are_rejected(mutate):
    frame = wire.encode_tensors([np.ones((2, 4), np.float32)])
    with pytest.raises(wire.WireError):
        wire.decode_tensors(mutate(frame))
    assert not wire.is_tensor_frame(b"{}") and not wire.is_tensor_frame({"x": 1})


def test_unsupported_inputs():
    with pytest.raises(wire.WireError):
        wire.encode_tensors([np.array(["a"])])
    with pytest.raises(wire.WireError):
        wire.encode_tensors([])
    with pytest.raises(wire.WireError):
        wire.encode_tensors([np.zeros((1,) * 9)])


def test_rest_binary_frames_in_and_out():
    """no user code: the frame IS the request; the reply comes back in the same framing; JSON bodies and gzip
    keep working on the same route; a malformed frame is the reference's 422"""
    from starlette.testclient import TestClient
    from clearml_serving_b200.main import create_app
    p = ModelRequestProcessor()
    ep = ModelEndpoint(engine_type="b200", serving_url="sum")
    p._endpoints["sum"] = ep
    p._engine_processor_lookup["sum"] = make_fake_engine(ep, FakeModel(n_features=4))
    client = TestClient(create_app(p), raise_server_exceptions=False)
    try:
        X = np.arange(12, dtype=np.float32).reshape(3, 4)
        body = wire.encode_tensors([X])
        r = client.post("/serve/sum", content=body, headers={"Content-Type": wire.MEDIA_TYPE})
        assert r.status_code == 200 and r.headers["content-type"] == wire.MEDIA_TYPE
        (y,) = wire.decode_tensors(r.content)
        assert y.dtype == np.float32 and np.array_equal(y, X.sum(axis=1))
        # gzip-compressed frame (GzipRoute, main.py:32-50)
        r = client.post("/serve/sum", content=gzip.compress(body),
                        headers={"Content-Type": wire.MEDIA_TYPE, "Content-Encoding": "gzip"})
        assert r.status_code == 200 and np.array_equal(wire.decode_tensors(r.content)[0], X.sum(axis=1))
        # float64 payload is cast to the model's input dtype like any other request (preprocess_service.py:393)
        r = client.post("/serve/sum", content=wire.encode_tensors([X.astype(np.float64)]),
                        headers={"Content-Type": wire.MEDIA_TYPE})
        assert r.status_code == 200 and np.array_equal(wire.decode_tensors(r.content)[0], X.sum(axis=1))
        # wrong tensor count / truncated frame -> 422 with the reference's detail prefix
        r = client.post("/serve/sum", content=wire.encode_tensors([X, X]), headers={"Content-Type": wire.MEDIA_TYPE})
        assert r.status_code == 422 and "processing request" in r.json()["detail"]
        r = client.post("/serve/sum", content=body[:-8], headers={"Content-Type": wire.MEDIA_TYPE})
        assert r.status_code == 422 and "payload truncated" in r.json()["detail"]
    finally:
        p.shutdown()


def test_frames_survive_arbitrary_tensors_and_reject_arbitrary_damage():
    """property test: any list of supported tensors round-trips bit for bit; any truncation / single-byte corruption of
    the header is either rejected with WireError or decodes to well-formed arrays (never an out-of-bounds view)"""
    from hypothesis import given, settings, strategies as st
    from hypothesis.extra import numpy as hnp
    dtypes = st.sampled_from([np.float32, np.float64, np.int32, np.int64, np.uint8, np.int8, np.bool_, np.uint64, np.float16,
                              np.uint32])
    arrays = dtypes.flatmap(lambda dt: hnp.arrays(dt, hnp.array_shapes(min_dims=0, max_dims=4, min_side=0, max_side=5)))

    @settings(max_examples=150, deadline=None)
    @given(st.lists(arrays, min_size=1, max_size=wire.MAX_TENSORS), st.data())
    def run(tensors, data):
        frame = wire.encode_tensors(tensors)
        back = wire.decode_tensors(frame)
        assert len(back) == len(tensors)
        for a, b in zip(tensors, back):
            assert a.dtype == b.dtype and a.shape == b.shape and a.tobytes() == b.tobytes()
        # damage: cut the frame anywhere, or flip one header byte
        cut = data.draw(st.integers(0, len(frame) - 1))
        pos = data.draw(st.integers(0, min(len(frame), 8 + 40 * len(tensors)) - 1))
        damaged = bytearray(frame)
        damaged[pos] ^= data.draw(st.integers(1, 255))
        for bad in (frame[:cut], bytes(damaged)):
            try:
                out = wire.decode_tensors(bad)
            except wire.WireError:
                continue
            total = sum(o.nbytes for o in out)
            assert total <= len(bad) and all(o.size == int(np.prod(o.shape)) for o in out)
    run()
