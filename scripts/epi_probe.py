"""Developer aid (ncu target): the epilogue-bound 1x1 expansion GEMM of ResNet-50 layer1 (M=401408, N=256, K=64), plain and
with the fp16 residual, a handful of launches each."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from clearml_serving_b200 import native  # noqa: E402

native.ensure_init(0)
lib = native.lib()
M, N, K = 401408, 256, 64
A = (torch.randn(M, K, device="cuda") * 0.5).half()
B = (torch.randn(N, K, device="cuda") * 0.05).half()
C = torch.empty(M, N, device="cuda", dtype=torch.half)
R = torch.randn(M, N, device="cuda").half()
bias = torch.randn(N, device="cuda")
for _ in range(3):
    native.check(lib.b2s_op_gemm(0, None, A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, None, None, 0, 0, 0))
    native.check(lib.b2s_op_gemm(0, None, A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, bias.data_ptr(), R.data_ptr(), 2, 0, 0))
torch.cuda.synchronize()
