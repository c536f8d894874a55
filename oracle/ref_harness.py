"""Run the reference's OWN hot-path code (from /root/reference) in this container, under stubs.

TEST INFRASTRUCTURE ONLY, and only usable where /root/reference exists (the build container).
It is used by oracle/gen_golden.py to record golden vectors into tests/golden/; nothing that runs
on the GPU box imports this module.

Stubs:
  * oracle/refstubs/clearml         -- fake `clearml` SDK (imports at model_request_processor.py:16-18,
                                       preprocess_service.py:11-13)
  * vllm.entrypoints.openai.protocol -- pydantic request types with `.model` (main.py:14,221,226)
  * clearml_serving.serving.init.setup_task -- returns (id, logger, instance) (main.py:21,59)
  * tritonclient.{grpc,utils,grpc.aio} -- in-process fake of the gRPC stub so that the reference's
    TritonPreprocessRequest.process (preprocess_service.py:313-446) runs unmodified against a
    Python callable standing in for tritonserver.
"""
import os
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_reference():
    """/root/reference in the build container; on the GPU box the pip --target install of it under baseline/_ref
    (git-ignored, travels with the repo snapshot: `python -m pip install --no-index --no-build-isolation --no-deps
    --target baseline/_ref /root/reference`, also run by __graft_entry__.build() when the source tree is present)"""
    for c in (os.environ.get("B2S_REFERENCE_ROOT"), "/root/reference",
              os.path.join(os.path.dirname(_HERE), "baseline", "_ref")):
        if c and os.path.isdir(os.path.join(c, "clearml_serving", "serving")):
            return c
    return "/root/reference"


REFERENCE_ROOT = _find_reference()


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "clearml_serving", "serving"))


class _Logger(object):
    def __init__(self):
        self.lines = []

    def report_text(self, msg, *a, **k):
        self.lines.append(str(msg))


# ------------------------------------------------------------------ fake tritonclient (in-process)
class _Contents(object):
    _fields = ("int_contents", "uint_contents", "int64_contents", "uint64_contents",
               "bool_contents", "fp32_contents", "fp64_contents", "bytes_contents")

    def __init__(self):
        for f in self._fields:
            setattr(self, f, [])


class _InferInputTensor(object):
    def __init__(self):
        self.name = ""
        self.datatype = ""
        self.shape = []
        self.contents = _Contents()


class _InferRequestedOutputTensor(object):
    def __init__(self):
        self.name = ""


class _ModelInferRequest(object):
    InferInputTensor = _InferInputTensor
    InferRequestedOutputTensor = _InferRequestedOutputTensor

    def __init__(self):
        self.model_name = ""
        self.model_version = ""
        self.inputs = []
        self.outputs = []


class _OutTensor(object):
    def __init__(self, name, shape, datatype):
        self.name, self.shape, self.datatype = name, list(shape), datatype


class _ModelInferResponse(object):
    def __init__(self):
        self.outputs = []
        self.raw_output_contents = []


_TRITON_TO_NP = {"BOOL": np.bool_, "INT8": np.int8, "INT16": np.int16, "INT32": np.int32,
                 "INT64": np.int64, "UINT8": np.uint8, "UINT16": np.uint16, "UINT32": np.uint32,
                 "UINT64": np.uint64, "FP16": np.float16, "FP32": np.float32, "FP64": np.float64}


def _np_to_triton_dtype(np_dtype):
    # restates tritonclient.utils.np_to_triton_dtype (third-party, public)
    for k, v in _TRITON_TO_NP.items():
        if np_dtype == v:
            return k
    if np_dtype == bool:
        return "BOOL"
    if np_dtype == np.object_ or np_dtype == np.bytes_:
        return "BYTES"
    return None


class FakeTritonServer(object):
    """models: dict model_name -> callable(list_of_np_inputs) -> list_of_np_outputs.
    Decodes the typed `contents` exactly as tritonserver would (typed repeated field -> tensor of
    `datatype` and `shape`), records the wire-level request for the golden file."""
    models = {}
    last_request = None

    class Stub(object):
        def __init__(self, channel):
            self.channel = channel

        async def ModelInfer(self, request, compression=None, timeout=None):
            FakeTritonServer.last_request = request
            fn = FakeTritonServer.models[request.model_name]
            ins = []
            for t in request.inputs:
                np_t = _TRITON_TO_NP[t.datatype]
                field = None
                for f in _Contents._fields:
                    if len(getattr(t.contents, f)):
                        field = f
                vals = getattr(t.contents, field) if field else []
                ins.append(np.array(vals, dtype=np_t).reshape(list(t.shape)))
            outs = fn(ins)
            resp = _ModelInferResponse()
            for name, o in zip([o.name for o in request.outputs], outs):
                o = np.ascontiguousarray(o)
                resp.outputs.append(_OutTensor(name, o.shape, _np_to_triton_dtype(o.dtype.type)))
                resp.raw_output_contents.append(o.tobytes())
            return resp


def _install_fake_tritonclient():
    if "tritonclient" in sys.modules:
        return
    tc = types.ModuleType("tritonclient")
    tc_grpc = types.ModuleType("tritonclient.grpc")
    tc_utils = types.ModuleType("tritonclient.utils")
    tc_aio = types.ModuleType("tritonclient.grpc.aio")
    fake_grpc = types.SimpleNamespace(
        aio=types.SimpleNamespace(insecure_channel=lambda addr, options=None: ("chan", addr)),
        Compression=types.SimpleNamespace(Gzip="gzip", NoCompression=None),
    )
    tc_grpc.grpc = fake_grpc
    tc_utils.np_to_triton_dtype = _np_to_triton_dtype
    tc_aio.service_pb2 = types.SimpleNamespace(ModelInferRequest=_ModelInferRequest)
    tc_aio.service_pb2_grpc = types.SimpleNamespace(GRPCInferenceServiceStub=FakeTritonServer.Stub)
    tc.grpc, tc.utils = tc_grpc, tc_utils
    tc_grpc.aio = tc_aio
    sys.modules.update({"tritonclient": tc, "tritonclient.grpc": tc_grpc,
                        "tritonclient.utils": tc_utils, "tritonclient.grpc.aio": tc_aio})


# ------------------------------------------------------------------ reference import under stubs
_LOADED = {}


def load_reference():
    """Returns a namespace with the reference's modules: ps (preprocess_service),
    mrp (model_request_processor), endpoints, main (FastAPI app module), logger."""
    if _LOADED:
        return types.SimpleNamespace(**_LOADED)
    if not available():
        raise RuntimeError("reference not present at {}".format(REFERENCE_ROOT))
    stubs = os.path.join(_HERE, "refstubs")
    for p in (REFERENCE_ROOT, stubs):
        if p not in sys.path:
            sys.path.insert(0, p)
    _install_fake_tritonclient()

    # vllm protocol stub (the installed vllm 0.22 no longer has this module)
    from pydantic import BaseModel

    class CompletionRequest(BaseModel):
        model: str = ""

    class ChatCompletionRequest(BaseModel):
        model: str = ""

    proto = types.ModuleType("vllm.entrypoints.openai.protocol")
    proto.CompletionRequest = CompletionRequest
    proto.ChatCompletionRequest = ChatCompletionRequest
    for name in ("vllm", "vllm.entrypoints", "vllm.entrypoints.openai"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    sys.modules["vllm.entrypoints.openai.protocol"] = proto

    logger = _Logger()
    init_mod = types.ModuleType("clearml_serving.serving.init")
    init_mod.setup_task = lambda *a, **k: ("stub-service", logger, "stub-instance")
    import clearml_serving.serving  # noqa
    sys.modules["clearml_serving.serving.init"] = init_mod

    import clearml_serving.serving.preprocess_service as ps
    import clearml_serving.serving.model_request_processor as mrp
    import clearml_serving.serving.endpoints as endpoints
    import clearml_serving.serving.main as main
    main.app.router.on_startup.clear()
    _LOADED.update(ps=ps, mrp=mrp, endpoints=endpoints, main=main, logger=logger)
    return types.SimpleNamespace(**_LOADED)


def make_processor(ref, endpoints_dict, engines=None):
    """Hand-built ModelRequestProcessor carrying exactly the attributes process_request /
    _process_request touch (model_request_processor.py:134-163)."""
    p = ref.mrp.ModelRequestProcessor.__new__(ref.mrp.ModelRequestProcessor)
    p._task = None
    p._metric_logging = {}
    p._kafka_stats_url = None
    p._metric_log_freq = 1.0
    p._stats_queue = ref.mrp.FastSimpleQueue()
    p._update_lock_flag = False
    p._canary_route = {}
    p._endpoints = dict(endpoints_dict)
    p._model_monitoring_endpoints = {}
    p._engine_processor_lookup = dict(engines or {})
    p._request_processing_state = ref.mrp.FastWriteCounter()
    p._enable_endpoint_telemetry = False
    p._endpoint_telemetry = {}
    return p


def make_engine(ref, engine_cls, endpoint, model=None, preprocess=None):
    """Engine object built without the clearml.Model fetch (skips preprocess_service.py:457)."""
    e = engine_cls.__new__(engine_cls)
    e.model_endpoint = endpoint
    e._preprocess = preprocess
    e._model = model
    e._timeout = 480
    if hasattr(engine_cls, "_ext_grpc"):
        # TritonPreprocessRequest.__init__ tail (preprocess_service.py:298-311)
        import tritonclient.grpc as tg
        import tritonclient.utils as tu
        import tritonclient.grpc.aio as ta
        e._ext_grpc = tg.grpc
        e._ext_np_to_triton_dtype = tu.np_to_triton_dtype
        e._ext_service_pb2 = ta.service_pb2
        e._ext_service_pb2_grpc = ta.service_pb2_grpc
        e._grpc_stub = {}
    return e
