"""b200 engine counterpart of the reference's examples/huggingface at BASELINE.json configs[3] (BERT-base, S <= 256): returns
the three id lists the reference's class returns (input_ids, token_type_ids, attention_mask); the engine packs the ragged
batch, nothing is padded.  Text bodies need a tokenizer folder on disk (`B2S_TOKENIZER_DIR`); token ids need nothing."""
import os
from typing import Any

MAX_LENGTH = 256


class Preprocess:
    def __init__(self):
        self.tokenizer = None
        folder = os.environ.get("B2S_TOKENIZER_DIR")
        if folder:
            from transformers import AutoTokenizer
            self.tokenizer = AutoTokenizer.from_pretrained(folder, local_files_only=True)

    def preprocess(self, body: dict, state: dict, collect_custom_statistics_fn=None) -> Any:
        if "input_ids" in body:
            ids = [int(t) for t in body["input_ids"]][:MAX_LENGTH]
            types = [int(t) for t in body.get("token_type_ids", [0] * len(ids))][:len(ids)]
            mask = [int(t) for t in body.get("attention_mask", [1] * len(ids))][:len(ids)]
        else:
            if self.tokenizer is None:
                raise ValueError("text requests need B2S_TOKENIZER_DIR (a local tokenizer folder); send input_ids instead")
            tok = self.tokenizer(text=body["text"], max_length=MAX_LENGTH, truncation=True)
            ids, mask = tok["input_ids"], tok["attention_mask"]
            types = tok.get("token_type_ids", [0] * len(ids))
        return [[ids], [types], [mask]]

    def postprocess(self, data: Any, state: dict, collect_custom_statistics_fn=None) -> dict:
        return {"data": data.tolist()}
