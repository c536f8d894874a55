"""User code of the LLM endpoint (engine `b200_llm`), the counterpart of the reference's examples/vllm/preprocess.py:
`load()` hands the engine what it cannot ship itself -- a tokenizer -- and the pre/post hooks see every OpenAI-route body.

The reference's example builds vLLM's serving objects in `load()` (examples/vllm/preprocess.py:13-60); here the engine
owns the model and the scheduler, so `load()` only returns `{"tokenizer": tok}`: any object with `encode`, `decode` and
(for chat) `apply_chat_template`.  Without network access no tokenizer files can be fetched: point B2S_TOKENIZER_PATH at
a local `transformers` tokenizer folder, or send token-id prompts (the OpenAI completions API allows them), which need
no tokenizer at all."""
import os
from typing import Any, Callable, Optional


class Preprocess(object):
    def load(self, local_file_name: Optional[str]) -> Any:
        path = os.environ.get("B2S_TOKENIZER_PATH") or local_file_name
        if path and os.path.isdir(path) and os.path.exists(os.path.join(path, "tokenizer_config.json")):
            from transformers import AutoTokenizer
            return {"tokenizer": AutoTokenizer.from_pretrained(path, local_files_only=True)}
        return {}          # token-id prompts only

    def preprocess(self, body: Any, state: dict, collect_custom_statistics_fn: Optional[Callable[[dict], None]] = None) -> Any:
        # bodies arrive as {"request": <OpenAI request>, ...} on the /openai/v1/* routes (serving/main.py:217-231)
        return body

    def postprocess(self, data: Any, state: dict, collect_custom_statistics_fn: Optional[Callable[[dict], None]] = None) -> Any:
        return data
