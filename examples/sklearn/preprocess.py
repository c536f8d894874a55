"""b200 engine counterpart of the reference's examples/sklearn: four iris features in, class label out."""
from typing import Any

import numpy as np


class Preprocess(object):
    FEATURES = ("x0", "x1", "x2", "x3")

    def preprocess(self, body: dict, state: dict, collect_custom_statistics_fn=None) -> Any:
        # one row per request; the engine batches rows of concurrent requests
        return [[float(body.get(k, 0.0)) for k in self.FEATURES]]

    def postprocess(self, data: Any, state: dict, collect_custom_statistics_fn=None) -> dict:
        labels = data[0] if isinstance(data, (list, tuple)) else data     # (labels, scores) for linear classifiers
        return dict(y=np.asarray(labels).tolist())
