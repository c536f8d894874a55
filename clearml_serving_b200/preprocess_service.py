"""Engine plugin layer: the reference's plugin ABI re-stated, plus the `b200` engine.

What is mirrored (names, arity, sync/async flags and error behaviour are the contract):
  * BasePreprocessRequest             clearml_serving/serving/preprocess_service.py:25-264
      - owns the user `Preprocess` object (`_preprocess`) and the model (`_model`)
      - preprocess/postprocess delegate to user code when it defines them, else pass through
        (:122-151, :153-180); `process` is the engine's job (:182-206)
      - `register_engine(name, modules=[...])` class registry (:230-243), `get_engine_cls` (:226-228),
        `validate_engine_type` (:222-224), `load_modules` (:245-253), server config (:214-220)
      - user code is loaded from the endpoint's `preprocess_artifact`, wrapped so that `unload()` runs
        when the engine is dropped, given a `send_request` helper, and its optional `load()` result
        becomes `_model` (:63-120)
  * the user plugin surface `class Preprocess` with optional load/unload/preprocess/process/
    postprocess/send_request  (clearml_serving/preprocess/preprocess_template.py:6-168)
  * TritonPreprocessRequest.process marshalling rules (:385-446) -- re-implemented by
    B200PreprocessRequest.process in front of libb200serve instead of a gRPC stub.

What is new: B200PreprocessRequest -- `is_process_async = True` like the Triton engine (:289-291);
`process` casts the request exactly like the reference client, enqueues it on the endpoint's
DynamicBatcher and awaits the batch result.  No CPU path exists: without the CUDA library/device
the constructor raises.
"""
import asyncio
import hashlib
import importlib
import importlib.util
import os
import sys
import threading
import traceback
from pathlib import Path

import numpy as np

from . import formats, model_repo, native, wire
from .router import Replica, ReplicaSet, parse_devices
from .scheduler import BatchPolicy, DynamicBatcher


def _default_timeout():
    # 80% of the web-server timeout, as preprocess_service.py:48-49
    return int(float(os.environ.get("GUNICORN_SERVING_TIMEOUT", 600)) * 0.8)


class BasePreprocessRequest(object):
    _engines = {}            # engine name -> class
    _engine_modules = set()  # best-effort pre-imports (before worker fork)
    _default_serving_base_url = "http://127.0.0.1:8080/serve/"
    _server_config = {}
    _model_resolver = None   # callable(model_id) -> local path ; defaults to clearml.Model if importable
    _timeout = None
    is_preprocess_async = False
    is_process_async = False
    is_postprocess_async = False

    def __init__(self, model_endpoint, task=None):
        # one object per endpoint per process, shared by all in-flight requests
        self.model_endpoint = model_endpoint
        self._preprocess = None
        self._model = None
        if self._timeout is None:
            self._timeout = _default_timeout()
        artifact = getattr(model_endpoint, "preprocess_artifact", None)
        if artifact:
            try:
                self._load_user_code(task, artifact)
            except Exception as ex:
                raise ValueError("Error: Failed loading preprocess code for '{}': {}\n\n{}".format(
                    artifact, ex, traceback.format_exc()))

    # ------------------------------------------------------------------ user code
    def _locate_user_code(self, task, artifact):
        if task is not None and artifact in getattr(task, "artifacts", {}):
            entry = task.artifacts[artifact]
            path = entry.get_local_copy(extract_archive=False)
            if not path or not Path(path).exists():
                raise ValueError("Artifact '{}' could not be downloaded".format(artifact))
            expected = getattr(entry, "hash", None)
            if expected:
                h = hashlib.sha256()
                with open(path, "rb") as f:
                    for chunk in iter(lambda: f.read(1 << 16), b""):
                        h.update(chunk)
                if h.hexdigest() != expected:
                    print("INFO: re-downloading artifact '{}' hash changed".format(artifact))
                    return entry.get_local_copy(extract_archive=True, force_download=True)
            return entry.get_local_copy(extract_archive=True)
        if task is None and Path(str(artifact)).exists():
            return str(artifact)  # standalone use: the artifact name is a local file / package dir
        raise ValueError("Error: could not find preprocessing artifact '{}' on Task id={}".format(
            artifact, getattr(task, "id", None)))

    def _load_user_code(self, task, artifact):
        path = Path(self._locate_user_code(task, artifact))
        if path.is_file():
            spec = importlib.util.spec_from_file_location("Preprocess", path.as_posix())
            module = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(module)
        else:
            name = str(artifact).replace(".", "_").replace("/", "_")
            spec = importlib.util.spec_from_file_location(
                name, location=(path / "__init__.py").as_posix(),
                submodule_search_locations=[path.as_posix()] + sys.path)
            module = importlib.util.module_from_spec(spec)
            sys.modules[spec.name] = module
            spec.loader.exec_module(module)
        user_cls = module.Preprocess

        class _Managed(user_cls):
            # dropping the engine (config change) must release what load() acquired
            def __del__(self):
                parent = super(_Managed, self)
                unload = getattr(parent, "unload", None)
                if callable(unload):
                    try:
                        unload()
                    except Exception as ex:  # noqa
                        print("Failed unloading model: {}".format(ex))
                fin = getattr(parent, "__del__", None)
                if callable(fin):
                    fin()

        _Managed.send_request = self._user_send_request()
        self._preprocess = _Managed()
        self._preprocess.model_endpoint = self.model_endpoint
        if callable(getattr(self._preprocess, "load", None)):
            self._model = self._preprocess.load(self._get_local_model_file())

    def _user_send_request(self):
        return BasePreprocessRequest._preprocess_send_request

    # ------------------------------------------------------------------ the three stages
    def preprocess(self, request, state, collect_custom_statistics_fn=None):
        """request body -> object handed to process(); raise to report an error."""
        if self._preprocess is not None and hasattr(self._preprocess, "preprocess"):
            return self._preprocess.preprocess(request, state, collect_custom_statistics_fn)
        return request

    def process(self, data, state, collect_custom_statistics_fn=None):
        """the model call; engines override."""
        return None

    def postprocess(self, data, state, collect_custom_statistics_fn=None):
        """model output -> dict returned by the REST layer."""
        if self._preprocess is not None and hasattr(self._preprocess, "postprocess"):
            return self._preprocess.postprocess(data, state, collect_custom_statistics_fn)
        return data

    # ------------------------------------------------------------------ model files
    def _get_local_model_file(self):
        model_id = getattr(self.model_endpoint, "model_id", None)
        if not model_id:
            return None
        resolver = BasePreprocessRequest._model_resolver
        self._model_framework = None   # the registry's framework tag, when known (triton_helper.py:159)
        if resolver is not None:
            r = resolver(model_id)
            if isinstance(r, tuple):   # (local_path, framework)
                r, self._model_framework = r[0], (r[1] if len(r) > 1 else None)
            return r
        if os.path.exists(str(model_id)):
            return str(model_id)
        try:
            from clearml import Model  # the control plane's model registry, when deployed with it
        except ImportError:
            raise ValueError("model '{}' is not a local file and no model resolver is configured".format(model_id))
        m = Model(model_id=model_id)
        self._model_framework = getattr(m, "framework", None)
        return m.get_local_copy()

    @classmethod
    def set_model_resolver(cls, resolver):
        BasePreprocessRequest._model_resolver = resolver

    # ------------------------------------------------------------------ registry / config
    @classmethod
    def set_server_config(cls, server_config):
        BasePreprocessRequest._server_config = server_config

    @classmethod
    def get_server_config(cls):
        return BasePreprocessRequest._server_config

    @classmethod
    def validate_engine_type(cls, engine):
        return engine in BasePreprocessRequest._engines

    @classmethod
    def get_engine_cls(cls, engine):
        return BasePreprocessRequest._engines.get(engine)

    @staticmethod
    def register_engine(engine_name, modules=None):
        def decorator(engine_cls):
            BasePreprocessRequest._engines[engine_name] = engine_cls
            return engine_cls
        if modules:
            BasePreprocessRequest._engine_modules.update(modules)
        return decorator

    @staticmethod
    def load_modules():
        for name in list(BasePreprocessRequest._engine_modules):
            try:
                importlib.import_module(name)
            except (ImportError, TypeError):
                pass  # best effort, like the reference

    @staticmethod
    def _serving_url(endpoint, version):
        endpoint = endpoint.strip("/")
        if version:
            endpoint = "{}/{}".format(endpoint, version.strip("/"))
        base = BasePreprocessRequest.get_server_config().get("base_serving_url") or \
            BasePreprocessRequest._default_serving_base_url
        return "{}/{}".format(base.strip("/"), endpoint)

    @staticmethod
    def _preprocess_send_request(_, endpoint, version=None, data=None):
        """`self.send_request(endpoint, version, data)` for user code: POST to a sibling endpoint,
        None on a non-2xx answer (preprocess_service.py:255-264)."""
        from requests import post
        reply = post(BasePreprocessRequest._serving_url(endpoint, version), json=data,
                     timeout=BasePreprocessRequest._timeout)
        return reply.json() if reply.ok else None


@BasePreprocessRequest.register_engine("custom")
class CustomPreprocessRequest(BasePreprocessRequest):
    """All three stages are user code, called synchronously (preprocess_service.py:504-517)."""

    def process(self, data, state, collect_custom_statistics_fn=None):
        if self._preprocess is not None and hasattr(self._preprocess, "process"):
            return self._preprocess.process(data, state, collect_custom_statistics_fn)
        return None


@BasePreprocessRequest.register_engine("custom_async")
class CustomAsyncPreprocessRequest(BasePreprocessRequest):
    """All three stages are user coroutines (preprocess_service.py:520-616)."""
    is_preprocess_async = True
    is_process_async = True
    is_postprocess_async = True

    def _user_send_request(self):
        return CustomAsyncPreprocessRequest._preprocess_send_request

    async def preprocess(self, request, state, collect_custom_statistics_fn=None):
        if self._preprocess is not None and hasattr(self._preprocess, "preprocess"):
            return await self._preprocess.preprocess(request, state, collect_custom_statistics_fn)
        return request

    async def process(self, data, state, collect_custom_statistics_fn=None):
        if self._preprocess is not None and hasattr(self._preprocess, "process"):
            return await self._preprocess.process(data, state, collect_custom_statistics_fn)
        return None

    async def postprocess(self, data, state, collect_custom_statistics_fn=None):
        if self._preprocess is not None and hasattr(self._preprocess, "postprocess"):
            return await self._preprocess.postprocess(data, state, collect_custom_statistics_fn)
        return data

    @staticmethod
    async def _preprocess_send_request(_, endpoint, version=None, data=None):
        from requests import post
        reply = await asyncio.to_thread(post, BasePreprocessRequest._serving_url(endpoint, version),
                                        json=data, timeout=BasePreprocessRequest._timeout)
        return reply.json() if reply.ok else None


# dtypes the reference's Triton client can put on the wire (_content_lookup, :271-282); fp16 is
# absent there too: fp16 is a compute dtype, never a request dtype.
_WIRE_DTYPES = (np.int32, np.uint8, np.int8, np.int64, np.uint64, np.uint32, np.bool_, np.float32,
                np.float64, int, bool)


class B200EngineMixin(object):
    """The engine logic, kept free of any base class so that it can be mixed into this package's
    BasePreprocessRequest AND into the reference's own base class (see integration.py)."""
    is_preprocess_async = False
    is_process_async = True
    is_postprocess_async = False
    _default_device = None

    def _b200_setup(self):
        ep = self.model_endpoint
        aux = getattr(ep, "auxiliary_cfg", None)
        packed = None
        if self._model is not None:
            # user load() returned something: accept a PackedModel, a path, or an sklearn estimator
            packed = self._pack_any(self._model)
        if packed is None:
            path = self._get_local_model_file()
            if not path:
                raise ValueError("b200 engine: endpoint '{}' has no model (model_id / load())".format(ep.serving_url))
            model_repo.validate_auxiliary_cfg(aux)
            packed, _fp = model_repo.default_repository().get(path, getattr(self, "_model_framework", None))
        self._packed_description = packed.description
        self._policy = BatchPolicy.from_auxiliary_cfg(aux)
        name = str(ep.serving_url).replace("/", "_")
        # variable-length models: per-row token capacity (aux `b200.max_seq_len`, else the model's own limit)
        self._max_row_elems = int(packed.description.get("max_row_elems", 0))
        if isinstance(aux, dict) and aux.get("b200.max_seq_len"):
            self._max_row_elems = min(self._max_row_elems or (1 << 30), int(aux["b200.max_seq_len"]))
        replicas = []
        for dev in parse_devices(aux, self._default_device_index()):
            m = native.Model(packed.kind, packed.blob, device=dev)   # one full model copy per GPU
            replicas.append(Replica(dev, m, DynamicBatcher(m, self._policy, name="{}@{}".format(name, dev),
                                                           max_row_elems=self._max_row_elems,
                                                           request_timeout_s=self._timeout)))
        self._replicas = ReplicaSet(replicas)
        self._native_model = replicas[0].model
        self._batcher = replicas[0].batcher
        self._model = self._native_model

    @staticmethod
    def _pack_any(obj):
        if isinstance(obj, formats.PackedModel):
            return obj
        if isinstance(obj, (str, os.PathLike)):
            return model_repo.load_model(str(obj))
        if isinstance(obj, dict) and "learner" in obj:
            return formats.pack_xgboost_json(obj)
        if hasattr(obj, "predict"):
            return formats.pack_sklearn(obj)
        if hasattr(obj, "state_dict") and hasattr(obj, "config"):   # a transformers torch module
            return formats.pack_torch_module(obj)
        return None

    def _default_device_index(self):
        if B200EngineMixin._default_device is not None:
            return int(B200EngineMixin._default_device)
        return int(os.environ.get("B2S_DEVICE", os.environ.get("LOCAL_RANK", 0)))

    def _next_batcher(self):
        reps = getattr(self, "_replicas", None)
        return reps.pick().batcher if reps is not None else self._batcher

    # ---- request marshalling: the rules of preprocess_service.py:385-406 --------------------------
    def _io_plan(self):
        """Per-engine constants of the marshalling rules, computed once: (single fixed-width input without an io spec ->
        its dtype and row width, else None; number of visible outputs; declared output dtypes or None)."""
        ep = self.model_endpoint
        m = self._native_model
        fast = None
        if not (getattr(ep, "input_name", None) and getattr(ep, "input_type", None)) and m.n_inputs == 1 \
                and m.in_row_elems[0] > 0:
            fast = (m.in_dtypes[0], m.in_row_elems[0])
        n_visible = len(getattr(ep, "output_name", None) or []) or (1 if m.info.kind != native.MODEL_GRAPH else m.n_outputs)
        n_visible = min(n_visible, m.n_outputs)
        types = getattr(ep, "output_type", None)
        out_types = [np.dtype(types[min(i, len(types) - 1)]) for i in range(n_visible)] if types else None
        self._b200_plan = (fast, n_visible, out_types)
        return self._b200_plan

    def _marshal(self, data):
        plan = getattr(self, "_b200_plan", None) or self._io_plan()
        fast = plan[0]
        if fast is not None and type(data) is np.ndarray and data.ndim == 2 and data.dtype == fast[0] \
                and data.shape[1] == fast[1] and data.flags.c_contiguous:
            return [data], data.shape[0]    # the common serving case: one [rows, features] array of the model's dtype
        ep = self.model_endpoint
        m = self._native_model
        names = getattr(ep, "input_name", None)
        types = getattr(ep, "input_type", None)
        if names and types:
            # single declared input keeps backwards compatibility: `data` IS the tensor
            list_data = [data] if len(names) == 1 else data
            sizes = getattr(ep, "input_size", None) or [None] * len(names)
            arrays = []
            for i_data, _name, m_type, _size in zip(list_data, names, types, sizes):
                dt = np.dtype(m_type)
                if dt.type not in _WIRE_DTYPES:
                    raise ValueError("Input type nt supported {}".format(dt.type))
                arrays.append(np.array(i_data, dtype=dt))
        else:
            # no io spec on the endpoint (as with the sklearn / xgboost engines): the model decides
            list_data = [data] if m.n_inputs == 1 else data
            arrays = [d if (isinstance(d, np.ndarray) and d.dtype == m.in_dtypes[i]) else np.array(d, dtype=m.in_dtypes[i])
                      for i, d in enumerate(list_data)]
        if len(arrays) != m.n_inputs:
            raise ValueError("b200 engine: model takes {} inputs, request carries {}".format(m.n_inputs, len(arrays)))
        rows = None
        out = []
        for i, a in enumerate(arrays):
            if a.dtype != m.in_dtypes[i]:
                a = a.astype(m.in_dtypes[i])
            re_ = m.in_row_elems[i]
            if a.ndim < 2:
                raise ValueError("Expected 2D array, got {}D array instead: a request carries its own "
                                 "batch dimension, e.g. [[x0, x1, ...]]".format(a.ndim))
            r = a.shape[0]
            a = np.ascontiguousarray(a).reshape(r, -1)
            if re_ < 0:   # variable-length row (token sequence)
                cap = getattr(self, "_max_row_elems", 0)
                if a.shape[1] == 0 or (cap and a.shape[1] > cap):
                    raise ValueError("b200 engine: sequence of {} tokens is outside (0, {}]".format(a.shape[1], cap))
                if out and a.shape[1] != out[0].shape[1]:
                    raise ValueError("b200 engine: inputs disagree on the sequence length")
            elif a.shape[1] != re_:
                raise ValueError("b200 engine: input {} has {} features per row, model expects {}".format(
                    i, a.shape[1], re_))
            if rows is not None and r != rows:
                raise ValueError("b200 engine: inputs disagree on the batch dimension")
            rows = r
            out.append(a)
        return out, rows

    def _unmarshal(self, outs):
        """np.frombuffer + np.resize semantics of preprocess_service.py:430-446: arrays own their
        memory, dtype from the endpoint's output_type (clamped to the last declared), single output
        returned bare."""
        _fast, n_visible, out_types = getattr(self, "_b200_plan", None) or self._io_plan()
        if n_visible == 1:
            return outs[0] if out_types is None else outs[0].astype(out_types[0], copy=False)
        if out_types is None:
            return list(outs[:n_visible])
        return [outs[i].astype(out_types[i], copy=False) for i in range(n_visible)]

    # ---- binary tensor frames (wire.py): a request body that IS a frame needs no user preprocess code, and the
    #      reply goes back in the same framing; JSON / user-defined bodies take the reference's path untouched
    def preprocess(self, request, state, collect_custom_statistics_fn=None):
        if (self._preprocess is None or not hasattr(self._preprocess, "preprocess")) and wire.is_tensor_frame(request):
            tensors = wire.decode_tensors(request)
            state["_b200_wire"] = True
            names = getattr(self.model_endpoint, "input_name", None)
            n_in = len(names) if names else self._native_model.n_inputs
            if len(tensors) != n_in:
                raise ValueError("b200 engine: frame carries {} tensors, endpoint takes {}".format(len(tensors), n_in))
            return tensors[0] if n_in == 1 else tensors
        return super(B200EngineMixin, self).preprocess(request, state, collect_custom_statistics_fn)

    def postprocess(self, data, state, collect_custom_statistics_fn=None):
        if state.get("_b200_wire") and (self._preprocess is None or not hasattr(self._preprocess, "postprocess")):
            from starlette.responses import Response
            outs = data if isinstance(data, (list, tuple)) else [data]
            return Response(content=wire.encode_tensors(outs), media_type=wire.MEDIA_TYPE)
        return super(B200EngineMixin, self).postprocess(data, state, collect_custom_statistics_fn)

    async def process(self, data, state, collect_custom_statistics_fn=None):
        # user override wins, exactly like the Triton engine (preprocess_service.py:340-341)
        if self._preprocess is not None and hasattr(self._preprocess, "process"):
            return await self._preprocess.process(data, state, collect_custom_statistics_fn)
        arrays, rows = self._marshal(data)
        # the deadline (self._timeout, preprocess_service.py:48-49) is enforced by the batcher on the queue age
        if collect_custom_statistics_fn is None:
            outs = await self._next_batcher().submit_async(arrays, rows)
        else:
            # a sampled request (model_request_processor.py:1341-1367) also reports what the batcher did with it: the
            # figures tritonserver kept per model behind :8002/metrics, here per request on the reference's stats channel
            info = {}
            outs = await self._next_batcher().submit_async(arrays, rows, info)
            collect_custom_statistics_fn({"_b200_batch_rows": info.get("batch_rows", 0),
                                          "_b200_queue_us": round(info.get("queue_us", 0.0), 1),
                                          "_b200_exec_us": round(info.get("exec_us", 0.0), 1)})
        return self._unmarshal(outs)

    def process_sync(self, data, timeout=None):
        """Blocking variant for non-asyncio callers (benchmarks, C++/thread hosts)."""
        arrays, rows = self._marshal(data)
        outs = self._next_batcher().submit(arrays, rows).result(timeout=timeout or self._timeout)
        return self._unmarshal(outs)

    def engine_stats(self):
        reps = getattr(self, "_replicas", None)
        return reps.snapshot_stats() if reps is not None else self._batcher.snapshot_stats()

    def unload(self):
        reps = getattr(self, "_replicas", None)
        if reps is not None:
            reps.shutdown()
            self._replicas = None
            self._batcher = None
            self._native_model = None
            return
        b = getattr(self, "_batcher", None)
        if b is not None:
            b.shutdown()
            self._batcher = None
        m = getattr(self, "_native_model", None)
        if m is not None:
            m.free()
            self._native_model = None

    def __del__(self):
        try:
            self.unload()
        except Exception:  # noqa
            pass


@BasePreprocessRequest.register_engine("b200", modules=["numpy"])
class B200PreprocessRequest(B200EngineMixin, BasePreprocessRequest):
    def __init__(self, model_endpoint, task=None):
        BasePreprocessRequest.__init__(self, model_endpoint=model_endpoint, task=task)
        self._b200_setup()
