"""Developer aid: attention_varlen on the BERT-base serving mix (64 sequences, S drawn from {16,64,128,256}, 12 heads)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from clearml_serving_b200 import native  # noqa: E402

native.ensure_init(0)
lib = native.lib()
rng = np.random.default_rng(1)
for name, lens in (("mix", rng.choice([16, 64, 128, 256], 64)), ("all256", np.full(64, 256)), ("all16", np.full(64, 16)),
                   ("all512", np.full(16, 512))):
    cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int64, device="cuda")
    T = int(lens.sum())
    qkv = (torch.randn(T, 3 * 768, device="cuda") * 0.5).half()
    out = torch.empty(T, 768, device="cuda", dtype=torch.half)
    def fn():
        native.check(lib.b2s_op_attention(0, None, qkv.data_ptr(), cu.data_ptr(), None, out.data_ptr(), len(lens), int(lens.max()), 12, 64, T))
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(50):
        fn()
    e.record(); torch.cuda.synchronize()
    us = s.elapsed_time(e) / 50 * 1e3
    flops = float((4.0 * lens.astype(np.float64) ** 2 * 64 * 12).sum())
    print("%-7s tokens=%5d  %.1f us  %.0f TFLOP/s (algorithmic 4*S^2*64 per head)" % (name, T, us, flops / us / 1e6))
