"""Request router: the dispatch half of the reference's ModelRequestProcessor, standalone.

Mirrors (behaviour, names, exceptions):
  * process_request          clearml_serving/serving/model_request_processor.py:253-304
      in-flight counter, stall-and-retry while a config swap is running, url normalisation, canary
      draw, endpoint lookup (EndpointNotFoundException), lazy engine construction + cache
  * _process_request         :1309-1369   preprocess -> <serve_type>() -> postprocess, awaited iff the
      engine's is_*_async flag says so; fresh `state` dict per request; latency + sampled stats
  * _process_canary          :306-313     numpy.random.choice over the route table
  * _normalize_endpoint_url  :1455-1457
  * FastWriteCounter         :58-70       lock-free in-flight counter
The control-plane half (ClearML task (de)serialisation, model auto-update, Kafka sender, plots) is
out of scope; endpoints are added programmatically / from a JSON file instead, with
the same swap protocol (:700-720) so that in-flight requests never see a half-updated table.
"""
import asyncio
import gc
import os
import itertools
import json
import threading
from collections import deque
from random import random
from time import sleep, time

from numpy.random import choice

from .endpoints import CanaryEP, ModelEndpoint
from .preprocess_service import BasePreprocessRequest


class ModelRequestProcessorException(Exception):
    pass


class EndpointNotFoundException(ModelRequestProcessorException):
    pass


class EndpointModelLoadException(ModelRequestProcessorException):
    pass


class EndpointBackendEngineException(ModelRequestProcessorException):
    pass


class ServingInitializationException(Exception):
    pass


class FastWriteCounter(object):
    """inc/dec without a lock: two monotonic itertools counters (next() is atomic under the GIL)."""

    def __init__(self):
        self._up = itertools.count()
        self._down = itertools.count()

    def inc(self):
        next(self._up)

    def dec(self):
        next(self._down)

    def value(self):
        return next(self._up) - next(self._down)


class ModelRequestProcessor(object):
    def __init__(self, task=None, name=None):
        self._task = task
        self._name = name or "b200-serving"
        self._endpoints = {}
        self._model_monitoring_endpoints = {}
        self._canary_endpoints = {}
        self._canary_route = {}
        self._engine_processor_lookup = {}
        self._metric_logging = {}
        self._metric_log_freq = 1.0
        self._stats_sink = None          # callable(dict) or None; stands in for the Kafka producer
        self._stats_queue = deque(maxlen=100000)
        self._update_lock_flag = False
        self._update_lock_guard = threading.Lock()
        self._request_processing_state = FastWriteCounter()

    def get_id(self):
        return self._name

    # ------------------------------------------------------------------ configuration
    def _swap(self, mutate):
        """Apply `mutate()` with no request in flight (model_request_processor.py:700-720), then drop
        all engines so they are rebuilt lazily against the new table (:1026-1028)."""
        with self._update_lock_guard:
            self._update_lock_flag = True
            try:
                while self._request_processing_state.value() != 0:
                    sleep(0.001)
                mutate()
                stale = self._engine_processor_lookup
                self._engine_processor_lookup = {}
            finally:
                self._update_lock_flag = False
        for eng in stale.values():
            unload = getattr(eng, "unload", None)
            if callable(unload):
                try:
                    unload()
                except Exception:  # noqa
                    pass
        del stale
        gc.collect()

    def add_endpoint(self, endpoint, preload=False):
        if isinstance(endpoint, dict):
            endpoint = ModelEndpoint(**endpoint)
        self._validate_model(endpoint)
        url = self._normalize_endpoint_url(endpoint.serving_url)

        def mutate():   # the canary table is rebuilt with every swap: a new version under a prefix is routed to at once
            self._endpoints[url] = endpoint
            self._update_canary_lookup()
        self._swap(mutate)
        if preload:
            self._get_engine(url, endpoint)
        return url

    def remove_endpoint(self, endpoint_url, version=None):
        url = self._normalize_endpoint_url(endpoint_url, version)
        if url not in self._endpoints:
            return False
        def mutate():   # ... and a removed endpoint loses its share of the canary traffic
            self._endpoints.pop(url, None)
            self._update_canary_lookup()
        self._swap(mutate)
        return True

    def add_canary_endpoint(self, canary):
        if isinstance(canary, dict):
            canary = CanaryEP(**canary)
        if not canary.load_endpoints and not canary.load_endpoint_prefix:
            raise ValueError("canary endpoint must have either load_endpoints or load_endpoint_prefix")

        def mutate():
            self._canary_endpoints[self._normalize_endpoint_url(canary.endpoint)] = canary
            self._update_canary_lookup()
        self._swap(mutate)

    def _update_canary_lookup(self):
        """Route table: public url -> {endpoints, weights}; weights normalised to sum 1, prefix
        matches sorted so that the newest version comes first (:772-814)."""
        table = {}
        for url, c in self._canary_endpoints.items():
            if c.load_endpoints:
                eps = [e for e in c.load_endpoints if e in self._endpoints]
                weights = [w for e, w in zip(c.load_endpoints, c.weights) if e in self._endpoints]
            else:
                prefix = c.load_endpoint_prefix
                eps = [e for e in self._endpoints if e.startswith(prefix)]

                def version_key(e):
                    tail = e[len(prefix):].strip("/")
                    return (0, -int(tail)) if tail.isdigit() else (1, tail)
                eps.sort(key=version_key)
                eps = eps[:len(c.weights)]
                weights = list(c.weights[:len(eps)])
            total = float(sum(weights))
            if not eps or total <= 0:
                continue
            table[url] = dict(endpoints=eps, weights=[w / total for w in weights])
        self._canary_route = table

    def load_endpoints_file(self, path):
        """JSON: {"endpoints": [ModelEndpoint dicts], "canary": [CanaryEP dicts]}."""
        with open(path, "rt") as f:
            cfg = json.load(f)
        for ep in cfg.get("endpoints", []):
            self.add_endpoint(ep)
        for c in cfg.get("canary", []):
            self.add_canary_endpoint(c)

    @classmethod
    def _validate_model(cls, endpoint):
        """The b200 engine needs no io description for tree / linear models (like the sklearn and
        xgboost engines); when the endpoint does declare one it must be complete, as for the Triton
        engine (model_request_processor.py:1518-1534)."""
        if endpoint.engine_type != "b200":
            return True
        d = endpoint.as_dict()
        io_keys = ["input_type", "input_size", "input_name", "output_type", "output_size", "output_name"]
        given = [k for k in io_keys if d.get(k)]
        if given and len(given) != len(io_keys):
            raise EndpointBackendEngineException(
                "b200 engine requires a complete input/output description - missing values in {}".format(
                    [k for k in io_keys if k not in given]))
        return True

    # ------------------------------------------------------------------ request path
    @classmethod
    def _normalize_endpoint_url(cls, endpoint, version=None):
        return "{}/{}".format(endpoint.rstrip("/"), version or "").rstrip("/")

    def _process_canary(self, base_url):
        route = self._canary_route.get(base_url)
        if not route:
            return None
        return choice(route["endpoints"], 1, p=route["weights"])[0]

    def _get_engine(self, url, ep):
        engine = self._engine_processor_lookup.get(url)
        if engine is None:
            engine_cls = BasePreprocessRequest.get_engine_cls(ep.engine_type)
            engine = engine_cls(model_endpoint=ep, task=self._task)
            self._engine_processor_lookup[url] = engine
        return engine

    async def process_request(self, base_url, version, request_body, serve_type="process"):
        self._request_processing_state.inc()
        if self._update_lock_flag:
            # a config swap is running: step out, wait, retry
            self._request_processing_state.dec()
            while self._update_lock_flag:
                await asyncio.sleep(0.5 + random())
            return await self.process_request(base_url=base_url, version=version,
                                              request_body=request_body, serve_type=serve_type)
        engine, url = None, None
        try:
            url = self._normalize_endpoint_url(base_url, version)
            routed = self._process_canary(base_url=url)
            if routed:
                url = routed
            ep = self._endpoints.get(url) or self._model_monitoring_endpoints.get(url)
            if not ep:
                raise EndpointNotFoundException("Model inference endpoint '{}' not found".format(url))
            engine = self._get_engine(url, ep)
            return await self._process_request(processor=engine, url=url, body=request_body, serve_type=serve_type)
        finally:
            if url and engine is not None and engine is not self._engine_processor_lookup.get(url):
                gc.collect()
            self._request_processing_state.dec()

    async def _process_request(self, processor, url, body, serve_type):
        sampled, custom_stats, stats_fn, freq = False, {}, None, 1
        metric_ep = self._metric_logging.get(url)
        if self._stats_sink is not None:
            freq = metric_ep.log_frequency if metric_ep is not None and \
                getattr(metric_ep, "log_frequency", None) is not None else self._metric_log_freq
            if freq and (freq >= 1 or random() <= freq):
                sampled, stats_fn = True, custom_stats.update
        tic = time()
        state = {}
        pre = processor.preprocess(body, state, stats_fn)
        if processor.is_preprocess_async:
            pre = await pre
        stage = getattr(processor, serve_type.replace("/", "_"))
        out = stage(pre, state, stats_fn)
        if processor.is_process_async:
            out = await out
        reply = processor.postprocess(out, state, stats_fn)
        if processor.is_postprocess_async:
            reply = await reply
        tic = time() - tic
        if sampled:
            stats = dict(_latency=round(tic, 4), _count=int(1.0 / freq), _url=url)
            stats.update(custom_stats)
            if metric_ep is not None:
                wanted = set(getattr(metric_ep, "metrics", {}).keys())
                if body and isinstance(body, dict):
                    stats.update({k: body[k] for k in set(body.keys()) & wanted})
                if reply and isinstance(reply, dict):
                    stats.update({k: reply[k] for k in set(reply.keys()) & wanted})
            try:
                self._stats_queue.append(stats)
                self._stats_sink(stats)
            except Exception:  # noqa
                pass
        return reply

    # ------------------------------------------------------------------ model hot reload
    def sync_models(self, repository=None):
        """One pass of the reference's periodic model sync, in process: `TritonHelper.model_service_update_step`
        (engines/triton/triton_helper.py:91-194, driven every `update_frequency_sec` by maintenance_daemon :226-289)
        re-fetches every endpoint's model and tritonserver (`--model-control-mode=poll`) reloads what changed.  Here
        the packed-model cache re-reads the model files; an endpoint whose model CONTENT or description changed gets
        its engine dropped with no request in flight -- only that endpoint's (the reference drops every engine on a
        configuration change, model_request_processor.py:1026-1028) -- and the next request rebuilds it on the GPU.
        Returns the list of reloaded urls."""
        from . import model_repo
        repo = repository or model_repo.default_repository()
        eps = dict(self._endpoints)
        eps.update(self._model_monitoring_endpoints)
        first = not getattr(self, "_models_synced", False)
        changed = repo.update_step(eps)
        self._models_synced = True
        stale_urls = [u for u in changed if u in self._engine_processor_lookup]
        if first or not stale_urls:
            return [] if first else stale_urls
        stale = []
        with self._update_lock_guard:
            self._update_lock_flag = True
            try:
                while self._request_processing_state.value() != 0:
                    sleep(0.001)
                for u in stale_urls:
                    eng = self._engine_processor_lookup.pop(u, None)
                    if eng is not None:
                        stale.append(eng)
            finally:
                self._update_lock_flag = False
        for eng in stale:
            unload = getattr(eng, "unload", None)
            if callable(unload):
                try:
                    unload()
                except Exception:  # noqa
                    pass
        del stale
        gc.collect()
        return stale_urls

    def start_sync_daemon(self, poll_frequency_sec=60.0, endpoints_file=None):
        """Background thread: every `poll_frequency_sec` (the reference's `--repository-poll-secs` /
        CLEARML_SERVING_POLL_FREQ) re-read `endpoints_file` when it changed and run sync_models()."""
        if getattr(self, "_sync_thread", None) is not None:
            return self._sync_thread
        self._sync_stop = threading.Event()
        state = dict(mtime=os.path.getmtime(endpoints_file) if endpoints_file and os.path.exists(endpoints_file) else None)

        def run():
            while not self._sync_stop.wait(poll_frequency_sec):
                try:
                    if endpoints_file and os.path.exists(endpoints_file):
                        mt = os.path.getmtime(endpoints_file)
                        if mt != state["mtime"]:
                            state["mtime"] = mt
                            self.load_endpoints_file(endpoints_file)
                    self.sync_models()
                except Exception as ex:  # noqa -- keep serving with what is loaded (the reference prints and carries on)
                    print("Warning: model sync failed: {}: {}".format(type(ex).__name__, ex))
        self._sync_thread = threading.Thread(target=run, name="b2s-model-sync", daemon=True)
        self._sync_thread.start()
        return self._sync_thread

    def stop_sync_daemon(self):
        t, self._sync_thread = getattr(self, "_sync_thread", None), None
        if t is not None:
            self._sync_stop.set()
            t.join(timeout=5)

    # ------------------------------------------------------------------ misc
    def set_stats_sink(self, sink, default_frequency=1.0):
        self._stats_sink = sink
        self._metric_log_freq = float(default_frequency)

    def engine_stats(self):
        return {url: e.engine_stats() for url, e in self._engine_processor_lookup.items() if hasattr(e, "engine_stats")}

    def on_request_endpoint_telemetry(self, base_url=None, version=None):
        pass

    def on_response_endpoint_telemetry(self, base_url=None, version=None):
        pass

    def shutdown(self):
        self.stop_sync_daemon()
        self._swap(lambda: None)
