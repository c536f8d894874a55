"""Developer aid: tcgen05 GEMM throughput on BERT-base / generic shapes vs torch (cuBLAS) on the same box."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from clearml_serving_b200 import native  # noqa: E402

native.ensure_init(0)
lib = native.lib()
torch.cuda.set_device(0)
shapes = [(8192, 768, 768), (8192, 2304, 768), (8192, 3072, 768), (8192, 768, 3072), (16384, 3072, 768),
          (4096, 4096, 4096), (8192, 8192, 8192), (512, 768, 768), (64, 768, 768)]
for (M, N, K) in shapes:
    A = (torch.randn(M, K, device="cuda") * 0.5).half()
    B = (torch.randn(N, K, device="cuda") * 0.05).half()
    C = torch.empty(M, N, device="cuda", dtype=torch.half)
    bias = torch.randn(N, device="cuda")
    def mine():
        native.check(lib.b2s_op_gemm(0, None, A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, bias.data_ptr(), None, 1, 0, 0))
    def mine_plain():
        native.check(lib.b2s_op_gemm(0, None, A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, None, None, 0, 0, 0))
    def mine_bias():
        native.check(lib.b2s_op_gemm(0, None, A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, bias.data_ptr(), None, 0, 0, 0))
    def ref():
        return torch.nn.functional.gelu(torch.nn.functional.linear(A, B, bias.half()))
    def plain():
        return torch.nn.functional.linear(A, B)
    res = {}
    for name, fn in (("b2s plain", mine_plain), ("b2s +bias", mine_bias), ("b2s +bias+gelu", mine), ("torch linear+gelu", ref), ("torch linear only", plain)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        s.record()
        for _ in range(n):
            fn()
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / n
        res[name] = (ms, 2.0 * M * N * K / ms / 1e9)
    r = ref().float(); mine(); torch.cuda.synchronize()
    err = (C.float() - r).abs().max().item() / r.abs().max().item()
    print("M=%d N=%d K=%d  " % (M, N, K) + "  ".join("%s: %.3f ms %.0f TF/s" % (k, v[0], v[1]) for k, v in res.items()) + "  relerr=%.2e" % err)
