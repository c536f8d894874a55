"""Developer aid: where the epilogue time of the K = 64 expansion GEMM goes (B2S_EPI_DBG probes) and the write floor."""
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


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3


print("dbg=%s" % os.environ.get("B2S_EPI_DBG", "0"),
      "plain %.1f us" % timeit(lambda: native.check(lib.b2s_op_gemm(0, None, A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, None, None, 0, 0, 0))),
      "residual %.1f us" % timeit(lambda: native.check(lib.b2s_op_gemm(0, None, A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, bias.data_ptr(), R.data_ptr(), 2, 0, 0))))
if os.environ.get("B2S_EPI_DBG", "0") == "0":
    print("write floor: zero_ 205MB %.1f us; copy_ 205MB->205MB %.1f us; add (2 reads 1 write) %.1f us" % (
        timeit(lambda: C.zero_()), timeit(lambda: C.copy_(R)), timeit(lambda: torch.add(C, R, out=C))))
