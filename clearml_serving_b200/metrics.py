"""Engine metrics in Prometheus text exposition.

The reference's Triton sidecar polls tritonserver's `:8002/metrics` and forwards every line of the form
`name{model="m",version="1"} value` as a scalar (clearml_serving/engines/triton/triton_helper.py:20, :45-89).  With the
b200 engine there is no tritonserver to scrape: `render()` produces the same lines -- tritonserver's `nv_inference_*`
names for the figures that exist in both worlds, `b200_*` for the rest -- from the per-endpoint batcher counters
(`B200EngineMixin.engine_stats`), and `main.create_app` serves them at `GET /metrics`.  Model names follow the Triton
client's rule `"{serving_url}_{version}"` without slashes (preprocess_service.py:375-379).
"""
from .scheduler import BATCH_ROWS_BUCKETS

# (metric, help, type, stats key or callable)
_COUNTERS = (
    ("nv_inference_request_success", "Number of successful inference requests", "counter",
     lambda s: s.get("requests", 0) - s.get("failed_requests", 0)),
    ("nv_inference_request_failure", "Number of failed inference requests", "counter", "failed_requests"),
    ("nv_inference_count", "Number of inferences performed (rows, does not include cached requests)", "counter", "rows"),
    ("nv_inference_exec_count", "Number of model executions performed (batches)", "counter", "batches"),
    ("nv_inference_queue_duration_us", "Cumulative inference queuing duration in microseconds", "counter",
     "queue_delay_us_sum"),
    ("nv_inference_compute_infer_duration_us", "Cumulative collate hand-off to results-on-host duration in microseconds",
     "counter", "exec_us_sum"),
    ("b200_input_bytes", "Request tensor bytes staged host to device", "counter", "in_bytes"),
    ("b200_output_bytes", "Result tensor bytes returned device to host", "counter", "out_bytes"),
    ("b200_max_batch_rows", "Largest batch dispatched so far", "gauge", "max_batch_rows"),
)


def triton_model_name(serving_url, version=None):
    name = "{}_{}".format(serving_url, version) if version else str(serving_url)
    return name.replace("/", "_").strip("_")


def _num(v):
    return ("%d" % v) if float(v).is_integer() else ("%.3f" % v)


def render(endpoint_stats):
    """{model name: (version string, stats dict from DynamicBatcher / ReplicaSet .snapshot_stats())} -> exposition text"""
    lines = []
    for metric, helptext, kind, key in _COUNTERS:
        lines.append("# HELP {} {}".format(metric, helptext))
        lines.append("# TYPE {} {}".format(metric, kind))
        for model, (version, st) in sorted(endpoint_stats.items()):
            value = key(st) if callable(key) else st.get(key, 0)
            lines.append('{}{{model="{}",version="{}"}} {}'.format(metric, model, version or "1", _num(value)))
    lines.append("# HELP b200_batch_rows Rows per dispatched batch")
    lines.append("# TYPE b200_batch_rows histogram")
    for model, (version, st) in sorted(endpoint_stats.items()):
        hist = st.get("batch_rows_hist") or []
        cum = 0
        for le, n in zip(list(BATCH_ROWS_BUCKETS) + ["+Inf"], hist):
            cum += n
            lines.append('b200_batch_rows_bucket{{model="{}",version="{}",le="{}"}} {}'.format(model, version or "1", le, cum))
        lines.append('b200_batch_rows_sum{{model="{}",version="{}"}} {}'.format(model, version or "1", _num(st.get("rows", 0))))
        lines.append('b200_batch_rows_count{{model="{}",version="{}"}} {}'.format(model, version or "1", _num(st.get("batches", 0))))
    return "\n".join(lines) + "\n"


def collect(processor):
    """engine statistics of every instantiated b200 endpoint of a ModelRequestProcessor"""
    out = {}
    for url, eng in list(getattr(processor, "_engine_processor_lookup", {}).items()):
        fn = getattr(eng, "engine_stats", None)
        if fn is None:
            continue
        try:
            st = fn()
        except Exception:  # noqa: an endpoint being torn down must not break the scrape
            continue
        ep = getattr(eng, "model_endpoint", None)
        version = str(getattr(ep, "version", "") or "")
        base = getattr(ep, "serving_url", None) or url
        out[triton_model_name(base, version)] = (version or "1", st)
    return out
