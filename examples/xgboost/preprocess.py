"""b200 engine counterpart of the reference's examples/xgboost at BASELINE.json configs[1] (32 features): the row goes to
the engine as a plain list -- no xgb.DMatrix -- and absent features are missing values, as the DMatrix would treat None."""
from typing import Any

import numpy as np

N_FEATURES = 32


class Preprocess(object):
    def preprocess(self, body: dict, state: dict, collect_custom_statistics_fn=None) -> Any:
        row = [body.get("x{}".format(i)) for i in range(N_FEATURES)]
        return np.array([[np.nan if v is None else float(v) for v in row]], dtype=np.float32)

    def postprocess(self, data: Any, state: dict, collect_custom_statistics_fn=None) -> dict:
        return dict(y=data.tolist() if isinstance(data, np.ndarray) else data)
