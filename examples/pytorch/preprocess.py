"""b200 engine counterpart of the reference's examples/pytorch at BASELINE.json configs[2] (ResNet-50, 3x224x224): encoded
image bytes are decoded and resized on the host, the uint8 pixels go to the engine as they are (the cast to fp16 happens on
the device, inside the stem's space-to-depth kernel); the reply is the arg-max class.  A body that already is a `B2ST`
tensor frame never reaches this class (the engine takes it directly, clearml_serving_b200/wire.py)."""
import io
from typing import Any, Union

import numpy as np


class Preprocess(object):
    SIZE = 224

    def preprocess(self, body: Union[bytes, dict], state: dict, collect_custom_statistics_fn=None) -> Any:
        if isinstance(body, (bytes, bytearray)):
            from PIL import Image
            try:
                image = Image.open(io.BytesIO(body)).convert("RGB").resize((self.SIZE, self.SIZE))
            except Exception:
                raise RuntimeError("Image could not be decoded")
            return np.ascontiguousarray(np.asarray(image, dtype=np.uint8).transpose(2, 0, 1)[None])   # [1, 3, H, W]
        pixels = np.asarray(body["pixels"], dtype=np.uint8)
        return pixels.reshape(-1, 3, self.SIZE, self.SIZE)

    def postprocess(self, data: Any, state: dict, collect_custom_statistics_fn=None) -> dict:
        logits = np.asarray(data)
        return {"class": [int(i) for i in logits.reshape(logits.shape[0], -1).argmax(axis=1)]}
