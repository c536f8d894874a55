"""Developer aid: per-item timeline (SM clock) of the tcgen05 attention kernel's first CTA.  B2S_ATTN_TIMING=1"""
import ctypes
import os
import sys

import numpy as np

os.environ["B2S_ATTN_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from clearml_serving_b200 import native  # noqa: E402

native.ensure_init(0)
lib = native.lib()
NAMES = ["prod:start", "prod:qk_empty", "mma:S start", "mma:qk_full", "mma:sfree", "mma:S issued", "mma:PV start", "mma:P ready", "mma:PV issued",
         "smx:start", "smx:S ready", "smx:P done", "smx:O ready", "smx:sfree"]
for name, lens in (("all16", np.full(64, 16)), ("all256", np.full(64, 256))):
    cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int64, device="cuda")
    T = int(lens.sum())
    qkv = (torch.randn(T, 3 * 768, device="cuda") * 0.5).half()
    out = torch.empty(T, 768, device="cuda", dtype=torch.half)
    for _ in range(3):
        native.check(lib.b2s_op_attention(0, None, qkv.data_ptr(), cu.data_ptr(), None, out.data_ptr(), len(lens), int(lens.max()), 12, 64, T))
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 256)()
    native.check(lib.b2s_debug_attention_stamps(buf))
    st = np.array(buf[:]).reshape(16, 16)
    t0 = st[0][0]
    print("==", name, "(cycles since the producer's first stamp)")
    for n in range(8):
        print(" item %d: " % n + "  ".join("%s=%d" % (NAMES[k].split(":")[1].replace(" ", "_"), st[n][k] - t0) for k in range(14)))
