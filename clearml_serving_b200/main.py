"""REST layer: the reference's endpoint contract in front of the b200 router.

Contract mirrored from clearml_serving/serving/main.py:
  * routes   POST {prefix}/{model_id} | {prefix}/{model_id}/ | {prefix}/{model_id}/{version}
             (main.py:191-193), prefix "/serve" or CLEARML_DEFAULT_SERVE_SUFFIX (:184);
             POST/GET {prefix}/openai/{endpoint_type:path} dispatching on serve_type (:217-231)
  * body     `Union[bytes, dict]` (:197); gzip-encoded bodies are inflated first (:32-50)
  * errors   EndpointNotFound -> 404, model-load / backend-engine -> 422, serving-initialisation -> 500,
             any other exception -> 422 unless its text carries the CUDA-OOM markers, in which case the
             worker exits so the supervisor restarts it (or answers 500 in dev mode) (:116-180)
  * reply    whatever postprocess returned, JSON-encoded by FastAPI (:205)
Detail strings are kept byte-identical (tests/golden/rest_contract.json was recorded from the
reference app).

Run:  B2S_ENDPOINTS=/path/endpoints.json uvicorn clearml_serving_b200.main:app
"""
import asyncio
import contextlib
import gzip
import os
import traceback
from http import HTTPStatus
from typing import Any, Callable, Dict, Optional, Union

from fastapi import APIRouter, Depends, FastAPI, HTTPException, Request, Response
from fastapi.responses import PlainTextResponse
from fastapi.routing import APIRoute
from starlette.background import BackgroundTask

from .model_request_processor import (
    EndpointBackendEngineException,
    EndpointModelLoadException,
    EndpointNotFoundException,
    ModelRequestProcessor,
    ServingInitializationException,
)

__version__ = "0.1.0"

_OOM_MARKERS = ("CUDA out of memory. ", "NVML_SUCCESS == r INTERNAL ASSERT FAILED")


class CUDAException(Exception):
    def __init__(self, exception):
        self.exception = exception


class _InflatingRequest(Request):
    async def body(self) -> bytes:
        if not hasattr(self, "_body"):
            raw = await super().body()
            if "gzip" in self.headers.getlist("Content-Encoding"):
                raw = gzip.decompress(raw)
            self._body = raw  # noqa
        return self._body


class GzipRoute(APIRoute):
    def get_route_handler(self) -> Callable:
        inner = super().get_route_handler()

        async def handler(request: Request) -> Response:
            return await inner(_InflatingRequest(request.scope, request.receive))

        return handler


class _Log(object):
    def report_text(self, msg, *_, **__):
        print(msg)


def create_app(processor: Optional[ModelRequestProcessor] = None, logger=None, instance_id="b200"):
    log = logger or _Log()
    state = dict(processor=processor)
    @contextlib.asynccontextmanager
    async def lifespan(_app):
        if state["processor"] is None:
            p = ModelRequestProcessor()
            cfg = os.environ.get("B2S_ENDPOINTS")
            if cfg:
                p.load_endpoints_file(cfg)
            # model hot reload (the reference polls every CLEARML_SERVING_POLL_FREQ minutes, serving/main.py:55-57;
            # its Triton sidecar every --repository-poll-secs): 0 disables
            poll_min = float(os.environ.get("B2S_POLL_FREQ_MIN", os.environ.get("CLEARML_SERVING_POLL_FREQ", "1.0")) or 0)
            if poll_min > 0:
                p.sync_models()
                p.start_sync_daemon(poll_min * 60.0, endpoints_file=cfg)
            state["processor"] = p
        yield
        if state["processor"] is not None:
            state["processor"].shutdown()

    app = FastAPI(title="ClearML Serving Service", version=__version__,
                  description="ClearML Service Service router (b200 engine)", lifespan=lifespan)
    app.state.b2s = state

    async def _stop_loop():
        asyncio.get_running_loop().stop()

    @app.exception_handler(CUDAException)
    async def _cuda_handler(request, exc):
        return PlainTextResponse("CUDA out of memory. Restarting service", status_code=500,
                                 background=BackgroundTask(_stop_loop))

    def _unprocessable(ex):
        text = str(ex)
        if any(m in text for m in _OOM_MARKERS):
            if os.environ.get("CLEARML_SERVING_DEV_CUDAEXCEPTION", "0") != "0":
                raise CUDAException(exception=ex)
            os._exit(1)  # cannot always recover: let the supervisor restart the worker
        raise HTTPException(status_code=422, detail="Error [{}] processing request: {}".format(type(ex), ex))

    def _report(kind, ex, request):
        log.report_text("[{}] Exception [{}] {} while {}: {}\n{}".format(
            instance_id, type(ex), ex, kind, request, "".join(traceback.format_exc())))

    async def process_with_exceptions(base_url, version, request, serve_type):
        p = state["processor"]
        p.on_request_endpoint_telemetry(base_url=base_url, version=version)
        try:
            reply = await p.process_request(base_url=base_url, version=version, request_body=request,
                                            serve_type=serve_type)
        except EndpointNotFoundException as ex:
            raise HTTPException(status_code=404,
                                detail="Error processing request, endpoint was not found: {}".format(ex))
        except (EndpointModelLoadException, EndpointBackendEngineException) as ex:
            _report("processing request", ex, request)
            raise HTTPException(status_code=422, detail="Error [{}] processing request: {}".format(type(ex), ex))
        except ServingInitializationException as ex:
            _report("loading serving inference", ex, request)
            raise HTTPException(status_code=500, detail="Error [{}] processing request: {}".format(type(ex), ex))
        except HTTPException:
            raise
        except Exception as ex:  # ValueError and everything else share the 422 / OOM path
            _report("processing request", ex, request)
            _unprocessable(ex)
        p.on_response_endpoint_telemetry(base_url=base_url, version=version)
        return reply

    router = APIRouter(
        prefix="/{}".format(os.environ.get("CLEARML_DEFAULT_SERVE_SUFFIX", "serve")),
        tags=["models"],
        responses={404: {"description": "Model Serving Endpoint Not found"}},
        route_class=GzipRoute,
    )

    @router.post("/{model_id}/{version}")
    @router.post("/{model_id}/")
    @router.post("/{model_id}")
    async def base_serve_model(model_id: str, version: Optional[str] = None,
                               request: Union[bytes, Dict[Any, Any]] = None):
        return await process_with_exceptions(base_url=model_id, version=version, request=request,
                                             serve_type="process")

    async def _json_only(raw_request: Request):
        media = raw_request.headers.get("content-type", "").lower().split(";", maxsplit=1)[0]
        if media != "application/json":
            raise HTTPException(status_code=HTTPStatus.UNSUPPORTED_MEDIA_TYPE,
                                detail="Unsupported Media Type: Only 'application/json' is allowed")

    @router.post("/openai/{endpoint_type:path}", dependencies=[Depends(_json_only)])
    @router.get("/openai/{endpoint_type:path}", dependencies=[Depends(_json_only)])
    async def openai_serve_model(endpoint_type: str, request: Dict[Any, Any], raw_request: Request):
        return await process_with_exceptions(base_url=request.get("model", ""), version=None,
                                             request={"request": request, "raw_request": raw_request},
                                             serve_type=endpoint_type)

    app.include_router(router)

    @app.get("/metrics", response_class=PlainTextResponse)
    async def engine_metrics():
        # what the reference's sidecar scraped from tritonserver :8002/metrics (triton_helper.py:45-89), same line format
        from . import metrics
        return PlainTextResponse(metrics.render(metrics.collect(state["processor"])),
                                 media_type="text/plain; version=0.0.4")

    return app


app = create_app()
