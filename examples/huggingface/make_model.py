"""BertForSequenceClassification with the bert-base configuration (12 layers, hidden 768, 2 labels), random weights, saved
with `save_pretrained` (config.json + safetensors): the folder the b200 engine loads.  `--tiny` for a 2-layer one."""
import os
import sys

import torch
from transformers import BertConfig, BertForSequenceClassification

args = [a for a in sys.argv[1:] if not a.startswith("--")]
out = os.path.join(args[0] if args else ".", "bert_model")
torch.manual_seed(0)
cfg = BertConfig(num_labels=2) if "--tiny" not in sys.argv else BertConfig(
    vocab_size=1000, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=512,
    max_position_embeddings=256, num_labels=2)
BertForSequenceClassification(cfg).eval().save_pretrained(out)
print(out)
