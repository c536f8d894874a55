"""Developer aid: the memory-bound residual GEMMs of ResNet-50 (1x1 expansions, batch 128) and BERT (fp32 residual stream):
achieved bytes/s of the algorithmic traffic (A + W + residual + C) against the HBM copy peak."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from clearml_serving_b200 import native  # noqa: E402

native.ensure_init(0)
lib = native.lib()
torch.cuda.set_device(0)
# (M, N, K, out_f32)
shapes = [(401408, 256, 64, 0), (100352, 512, 128, 0), (25088, 1024, 256, 0), (6272, 2048, 512, 0),
          (401408, 64, 256, 0), (7424, 768, 768, 1), (7424, 768, 3072, 1)]
for (M, N, K, f32) in shapes:
    A = (torch.randn(M, K, device="cuda") * 0.5).half()
    B = (torch.randn(N, K, device="cuda") * 0.05).half()
    odt = torch.float32 if f32 else torch.half
    C = torch.empty(M, N, device="cuda", dtype=odt)
    R = torch.randn(M, N, device="cuda").to(odt)
    bias = torch.randn(N, device="cuda")
    es = 4 if f32 else 2
    variants = {"plain": (None, 0), "+bias+relu": (None, 2), "+bias+residual": (R, 0)}
    out = []
    for name, (res, act) in variants.items():
        def fn():
            native.check(lib.b2s_op_gemm(0, None, A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K,
                                         None if name == "plain" else bias.data_ptr(), res.data_ptr() if res is not None else None,
                                         act, 0, f32))
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
        byts = M * K * 2 + N * K * 2 + M * N * es * (2 if res is not None else 1)
        out.append("%s: %.1f us %.0f TF/s %.2f TB/s" % (name, ms * 1e3, 2.0 * M * N * K / ms / 1e9, byts / ms / 1e9))
    ref = torch.nn.functional.linear(A, B).float() + bias + R.float()
    err = (C.float() - ref).abs().max().item() / ref.abs().max().item()
    print("M=%d N=%d K=%d %s  " % (M, N, K, "f32" if f32 else "f16") + "  ".join(out) + "  relerr=%.1e" % err)
