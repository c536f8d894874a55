"""one configuration of the attention op, a few launches (for ncu)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clearml_serving_b200 import native
native.ensure_init(0)
lib = native.lib()
lens = np.full(64, 256)
cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int64, device="cuda")
T = int(lens.sum())
qkv = (torch.randn(T, 3 * 768, device="cuda") * 0.5).half()
out = torch.empty(T, 768, device="cuda", dtype=torch.half)
for _ in range(4):
    native.check(lib.b2s_op_attention(0, None, qkv.data_ptr(), cu.data_ptr(), None, out.data_ptr(), len(lens), 256, 12, 64, T))
torch.cuda.synchronize()
