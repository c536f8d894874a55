"""The drop-in boundary, proven against the REFERENCE'S OWN dispatcher (CPU; needs the reference importable:
/root/reference in the build container or the baseline/_ref install on the GPU box).

`integration.register_with_reference("b200")` adds the engine to the reference's registry
(clearml_serving/serving/preprocess_service.py:230-243); the reference's ModelRequestProcessor.process_request
(model_request_processor.py:253-304) then builds the engine lazily from its ModelEndpoint (:287-291, model file through
the reference's own _get_local_model_file, preprocess_service.py:208-212) and runs its 3-stage pipeline (:1309-1369) over
it.  Only the native layer is faked (host-side model / stream executing the packed blob with the kernel's rules), so
everything above the C ABI -- the mixin inside the reference's class hierarchy, marshalling, batcher, futures -- is the
real code, driven by the real reference."""
import asyncio
import json

import numpy as np
import pytest

from oracle import oracle as orc
from oracle import ref_harness as rh
from tests import blob_interp
from tests.fakes import FakeStream

pytestmark = pytest.mark.skipif(not rh.available(), reason="reference package not importable here")


class _BlobModel(object):
    """native.Model stand-in: executes the packed forest blob on the host (tests/blob_interp.py)"""

    def __init__(self, kind, blob, device=0):
        d = blob_interp.decode(blob)
        self.kind, self.blob, self.device = kind, blob, device
        self.n_inputs, self.n_outputs = 1, 1
        self.in_dtypes = [np.dtype(np.float32)]
        self.out_dtypes = [np.dtype(np.float64 if d["acc_mode"] == 1 else np.float32)]
        self.in_row_elems, self.out_row_elems = [d["n_features"]], [1]
        self.fn = lambda x: blob_interp.predict(blob, x)

        class _I(object):
            kind = 1
        self.info = _I()
        self.freed = False

    def free(self):
        self.freed = True


@pytest.fixture()
def fake_native(monkeypatch):
    from clearml_serving_b200 import native, scheduler
    made = []

    def model(kind, blob, device=0):
        m = _BlobModel(kind, blob, device)
        made.append(m)
        return m
    monkeypatch.setattr(native, "Model", model)
    monkeypatch.setattr(scheduler.native, "Stream",
                        lambda m, max_rows, max_row_elems=0, n_slots=4: FakeStream(m, max_rows, n_slots=n_slots))
    return made


@pytest.mark.parametrize("engine_name", ["b200", "xgboost"])
def test_reference_dispatcher_serves_through_the_registered_b200_engine(tmp_path, fake_native, engine_name):
    import clearml_serving_b200.integration as b2s
    ref = rh.load_reference()
    import clearml   # the stub package (oracle/refstubs): Model(model_id).get_local_copy() -> a local path
    cls = b2s.register_with_reference(engine_name)          # "xgboost": shadow the built-in engine name
    assert ref.ps.BasePreprocessRequest.get_engine_cls(engine_name) is cls
    assert issubclass(cls, ref.ps.BasePreprocessRequest) and cls.is_process_async

    forest = orc.synth_xgb_forest(n_trees=31, depth=5, n_features=8, seed=12, ragged=True)
    path = tmp_path / "model.json"
    path.write_text(json.dumps(orc.xgb_json_from_forest(forest, base_score=0.5)))
    clearml.Model._paths["model-123"] = str(path)

    # user code exactly as the reference loads it: a Preprocess class from a task artifact (here injected after the
    # constructor ran, as oracle/ref_harness.make_engine does for the reference's own engines)
    class Pre(object):
        def preprocess(self, body, state, collect_custom_statistics_fn=None):
            return np.array([[body["x{}".format(i)] for i in range(8)]], dtype=np.float32)

        def postprocess(self, data, state, collect_custom_statistics_fn=None):
            return dict(y=data.tolist())
    ep = ref.endpoints.ModelEndpoint(engine_type=engine_name, serving_url="trees", model_id="model-123",
                                     auxiliary_cfg={"max_batch_size": 16, "dynamic_batching.max_queue_delay_microseconds": 2000})
    proc = rh.make_processor(ref, {"trees": ep})             # NO engine injected: process_request must build it
    rng = np.random.default_rng(0)
    X = rng.standard_normal((40, 8)).astype(np.float32)
    want = orc.forest_predict_xgb(forest, X, 0.5)

    async def run():
        first = await proc.process_request(base_url="trees", version=None,
                                           request_body=X[0:1].tolist(), serve_type="process")   # no user code yet: raw rows in
        eng = proc._engine_processor_lookup["trees"]
        assert type(eng) is cls and len(fake_native) == 1
        eng._preprocess = Pre()
        replies = await asyncio.gather(*[
            proc.process_request(base_url="trees", version=None, serve_type="process",
                                 request_body={"x{}".format(j): float(X[i, j]) for j in range(8)}) for i in range(40)])
        return first, replies, eng
    first, replies, eng = asyncio.run(run())
    assert np.float32(np.asarray(first).ravel()[0]) == want[0]
    got = np.array([r["y"][0] for r in replies], dtype=np.float32)
    assert np.array_equal(got, want)                                        # bit-exact through the reference's pipeline
    st = eng.engine_stats()
    assert st["requests"] == 41 and st["batches"] < 41                      # the 40 concurrent requests were batched
    # unknown endpoint: the reference's own exception type (-> 404 in its REST layer)
    with pytest.raises(ref.mrp.EndpointNotFoundException):
        asyncio.run(proc.process_request(base_url="nope", version=None, request_body={}, serve_type="process"))
    # engines are dropped on reconfiguration (model_request_processor.py:1026-1028): unload releases the native objects
    proc._engine_processor_lookup.clear()
    eng.unload()
    assert fake_native[0].freed
    clearml.Model._paths.pop("model-123", None)
