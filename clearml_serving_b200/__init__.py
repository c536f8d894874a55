"""b200serve: a B200-native (sm_100a) dynamic-batching inference engine behind the clearml-serving
engine-plugin API.  Importing the package registers the `b200` engine (plus the user-code engines
`custom` / `custom_async`) in this package's registry; `integration.register_with_reference()` does
the same inside an installed clearml-serving.

Layout (only what the hot path needs):
  csrc/                  CUDA kernels + the C ABI (include/b200serve.h) -> libb200serve.so
  native.py              ctypes binding (no fallback: raises when the library / GPU is missing)
  formats.py             model files the reference loads -> packed device blobs
  scheduler.py           per-endpoint queue + dynamic batcher (Triton semantics)
  preprocess_service.py  engine plugin ABI mirror + B200PreprocessRequest
  model_request_processor.py / endpoints.py / main.py   router, schema, REST contract mirror
  router.py              multi-GPU replica round-robin
"""
from .preprocess_service import (  # noqa: F401
    B200PreprocessRequest,
    BasePreprocessRequest,
    CustomAsyncPreprocessRequest,
    CustomPreprocessRequest,
)
from .endpoints import CanaryEP, ModelEndpoint  # noqa: F401

__version__ = "0.1.0"
