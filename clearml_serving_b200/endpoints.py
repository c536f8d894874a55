"""Endpoint records read by the b200 engine and its router.

The FIELD NAMES are the contract: they are the ones the reference's router and engines read
(clearml_serving/serving/endpoints.py:64-88 -- ModelEndpoint, CanaryEP), so an endpoint definition
written for the reference (`clearml-serving model add --engine ... --input-size ... --aux-config
...`, stored as a dict on the control-plane task) can be splatted into these records unchanged.
Normalisation rules restated from endpoints.py:11-33: scalar io fields become one-element lists, a
flat `input_size` list becomes a list of one shape, dtypes must be numpy-parsable, and the engine
name must be registered.
"""
from dataclasses import asdict, dataclass, field
from typing import Any, List, Optional

import numpy as np


def _as_list(value):
    if value is None:
        return None
    return list(value) if isinstance(value, (list, tuple)) else [value]


def _as_shape_list(value):
    """[-1, 2] -> [[-1, 2]];  [[3], [3]] stays;  7 -> [7]."""
    if value is None:
        return None
    if isinstance(value, (list, tuple)):
        if not any(isinstance(v, (list, tuple)) for v in value):
            return [list(value)]
        return list(value)
    return [value]


def _check_dtypes(values, what):
    for v in values or []:
        if not v:
            continue
        try:
            np.dtype(v)
        except TypeError:
            raise TypeError("{} not supported matrix type".format(v))


class _Record(object):
    def as_dict(self, remove_null_entries=False):
        d = asdict(self)
        if remove_null_entries:
            d = {k: v for k, v in d.items() if v is not None}
        return d


@dataclass
class ModelEndpoint(_Record):
    engine_type: str
    serving_url: str                       # full serving url incl. version, e.g. "detect_cat/v1"
    model_id: Optional[str] = None         # model to fetch (or a local path, see preprocess_service)
    version: str = ""
    preprocess_artifact: Optional[str] = None
    input_size: Optional[List[Any]] = None
    input_type: Optional[List[Any]] = None
    input_name: Optional[List[Any]] = None
    output_size: Optional[List[Any]] = None
    output_type: Optional[List[Any]] = None
    output_name: Optional[List[Any]] = None
    auxiliary_cfg: Any = None              # dict or pbtxt text: max_batch_size / dynamic_batching.*

    def __post_init__(self):
        from .preprocess_service import BasePreprocessRequest
        if not BasePreprocessRequest.validate_engine_type(self.engine_type):
            raise TypeError("{} not supported engine type".format(self.engine_type))
        self.input_size = _as_shape_list(self.input_size)
        self.output_size = _as_shape_list(self.output_size)
        self.input_type = _as_list(self.input_type)
        self.output_type = _as_list(self.output_type)
        self.input_name = _as_list(self.input_name)
        self.output_name = _as_list(self.output_name)
        _check_dtypes(self.input_type, "input_type")
        _check_dtypes(self.output_type, "output_type")


@dataclass
class CanaryEP(_Record):
    endpoint: str                           # the public (load-balanced) url
    weights: List[float]                    # one weight per routed endpoint
    load_endpoints: List[str] = field(default_factory=list)
    load_endpoint_prefix: Optional[str] = None
