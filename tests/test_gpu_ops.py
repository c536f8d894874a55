"""Operator-level GPU tests (-m gpu) for the DL building blocks, through the C ABI (b2s_op_*), each
against a plain fp32 reference of the same op (numpy / torch CPU).  Floating point: tolerance stated
per test (fp16 inputs, fp32 accumulation, fp16 or fp32 output)."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gemm(native, A, B, bias=None, residual=None, act=0, out_f32=False, bf16=False):
    M, K = A.shape
    N = B.shape[0]
    dA, dB = native.DeviceBuffer(A.nbytes), native.DeviceBuffer(B.nbytes)
    dC = native.DeviceBuffer(M * N * (4 if out_f32 else 2))
    dA.upload(A); dB.upload(B)
    dbias = dres = None
    if bias is not None:
        dbias = native.DeviceBuffer(bias.nbytes); dbias.upload(bias)
    if residual is not None:
        dres = native.DeviceBuffer(residual.nbytes); dres.upload(residual)
    try:
        native.check(native.lib().b2s_op_gemm(0, None, dA.ptr, dB.ptr, dC.ptr, M, N, K,
                                              dbias.ptr if dbias else None, dres.ptr if dres else None,
                                              act, 1 if bf16 else 0, 1 if out_f32 else 0))
        return dC.download(np.float32 if out_f32 else np.float16, M * N).reshape(M, N)
    finally:
        for b in (dA, dB, dC, dbias, dres):
            if b is not None:
                b.free()


def _ref_act(x, act):
    if act == 1:
        from scipy.special import erf
        return 0.5 * x * (1.0 + erf(x / np.sqrt(2.0)))
    if act == 2:
        return np.maximum(x, 0)
    if act == 3:
        return np.tanh(x)
    return x


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 128, 768), (256, 768, 768), (300, 2304, 768),
                                    (1000, 3072, 768), (77, 768, 3072), (5, 2, 768), (130, 200, 72)])
def test_gemm_fp16_matches_fp32_reference(gpu_native, M, N, K):
    rng = np.random.default_rng(M * 7 + N)
    A = (rng.standard_normal((M, K)) * 0.5).astype(np.float16)
    B = (rng.standard_normal((N, K)) * 0.05).astype(np.float16)
    ref = A.astype(np.float32) @ B.astype(np.float32).T
    got = _gemm(gpu_native, A, B, out_f32=True)
    # fp32 accumulation of exact fp16 products: only the summation order differs
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-4 * np.abs(ref).max())
    got16 = _gemm(gpu_native, A, B).astype(np.float32)
    np.testing.assert_allclose(got16, ref, rtol=2e-3, atol=2e-3 * np.abs(ref).max())


@pytest.mark.parametrize("act", [0, 1, 2, 3])
def test_gemm_fused_epilogue(gpu_native, act):
    rng = np.random.default_rng(act)
    M, N, K = 200, 768, 768
    A = (rng.standard_normal((M, K)) * 0.5).astype(np.float16)
    B = (rng.standard_normal((N, K)) * 0.05).astype(np.float16)
    bias = rng.standard_normal(N).astype(np.float32)
    res = rng.standard_normal((M, N)).astype(np.float16)
    ref = _ref_act(A.astype(np.float32) @ B.astype(np.float32).T + bias, act) + res.astype(np.float32)
    got = _gemm(gpu_native, A, B, bias=bias, residual=res, act=act).astype(np.float32)
    np.testing.assert_allclose(got, ref, rtol=2e-3, atol=2e-3 * np.abs(ref).max())
