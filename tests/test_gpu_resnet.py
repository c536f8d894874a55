"""GPU parity (-m gpu) for the convolutional graph path (BASELINE.json configs[2]: ResNet-50 fp16, 3x224x224)
against the torch-CPU-fp32 forward of the same weights (stand-in for the reference's Triton-CPU libtorch
backend; parity unpinned by reference tests).
Tolerance: max |logit - ref| / max |ref| <= 1e-3 (north_star), written below."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REL_TOL = 1e-3


def _realistic_bn(model, seed=0):
    """fresh BatchNorm layers are identities in eval mode, which lets activations grow by 2x per block;
    give them trained-looking statistics so that values stay O(1) like in a real checkpoint"""
    import torch
    g = torch.Generator().manual_seed(seed)
    for name, m in model.named_modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            with torch.no_grad():
                m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
                scale = 0.3 if name.endswith(("bn3", "bn2")) and "layer" in name else 1.0
                m.weight.copy_((torch.rand(m.num_features, generator=g) * 0.5 + 0.75) * scale)
                m.bias.copy_(torch.randn(m.num_features, generator=g) * 0.1)
    return model.eval()


def _run(native, pm, x, max_rows):
    model = native.Model(pm.kind, pm.blob, device=0)
    st = native.Stream(model, max_rows, 0, 1)
    try:
        reqs = [[x[i:i + 1]] for i in range(x.shape[0])]
        ev, outs, keep = st.infer_batch(reqs)
        st.wait(ev)
        got = np.concatenate([o[0] for o in outs])
        ev, outs2, keep = st.infer_batch([[x[1:3]]])     # a 2-image request, different batch composition
        st.wait(ev)
        assert np.array_equal(outs2[0][0], got[1:3]), "result depends on batch composition"
        return got
    finally:
        st.destroy()
        model.free()


@pytest.mark.parametrize("implicit", [True, False])
def test_resnet18_small_images(gpu_native, implicit):
    """implicit: KxK / strided convolutions as implicit GEMM (im2col-mode TMA); explicit: im2col kernel + GEMM"""
    import torch
    import torchvision
    from clearml_serving_b200 import formats
    torch.manual_seed(0)
    m = _realistic_bn(torchvision.models.resnet18(weights=None, num_classes=16))
    x = np.random.default_rng(1).standard_normal((5, 3, 64, 64)).astype(np.float32)
    with torch.no_grad():
        ref = m(torch.from_numpy(x)).numpy()
    got = _run(gpu_native, formats.pack_resnet(m, image_hw=(64, 64), implicit_conv=implicit), x, 8)
    assert got.shape == ref.shape and got.dtype == np.float32
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err <= REL_TOL, "relative error {:.2e}".format(err)


def test_resnet50_224(gpu_native):
    """configs[2]: torchvision resnet50 (25.6 M params), 3x224x224 fp32 images in, 1000 fp32 logits out."""
    import torch
    import torchvision
    from clearml_serving_b200 import formats
    torch.manual_seed(0)
    m = _realistic_bn(torchvision.models.resnet50(weights=None))
    x = np.random.default_rng(2).standard_normal((4, 3, 224, 224)).astype(np.float32)
    with torch.no_grad():
        ref = m(torch.from_numpy(x)).numpy()
    got = _run(gpu_native, formats.pack_resnet(m), x, 8)
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err <= REL_TOL, "relative error {:.2e}".format(err)


def test_resnet50_at_the_baseline_batch(gpu_native):
    """BASELINE.json configs[2] at its FULL batch: 128 single-image requests in one batch, every row against the
    torch-CPU-fp32 forward of the same weights"""
    import torch
    import torchvision
    from clearml_serving_b200 import formats
    torch.manual_seed(0)
    m = _realistic_bn(torchvision.models.resnet50(weights=None))
    x = np.random.default_rng(5).standard_normal((128, 3, 224, 224)).astype(np.float32)
    with torch.no_grad():
        ref = np.concatenate([m(torch.from_numpy(x[i:i + 16])).numpy() for i in range(0, 128, 16)])
    got = _run(gpu_native, formats.pack_resnet(m), x, 128)
    assert got.shape == (128, 1000)
    err = np.abs(got - ref).max(axis=1) / np.abs(ref).max()
    assert err.max() <= REL_TOL, "worst image {}: relative error {:.2e}".format(int(err.argmax()), err.max())


def test_resnet_uint8_input(gpu_native):
    import torch
    import torchvision
    from clearml_serving_b200 import formats
    torch.manual_seed(1)
    m = _realistic_bn(torchvision.models.resnet18(weights=None, num_classes=8), seed=3)
    x = np.random.default_rng(3).integers(0, 256, (3, 3, 64, 64)).astype(np.uint8)
    with torch.no_grad():
        ref = m(torch.from_numpy(x.astype(np.float32))).numpy()
    got = _run(gpu_native, formats.pack_resnet(m, input_dtype="uint8", image_hw=(64, 64)), x, 4)
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err <= 2e-3, "relative error {:.2e} (uint8 pixels up to 255 feed fp16 activations)".format(err)


def test_large_batch_collate_pool_matches_small_batches(gpu_native):
    """b2s_infer_batch collates batches >= 2 MB with the worker pool, group by group, with the host->device copy of a
    group enqueued while the next one is gathered (csrc/api.cu): a 41-request batch of 128x128 fp32 images (15.7 MB;
    requests of 1, 2 and 3 images, so group boundaries fall inside and between requests) must give exactly the rows the
    same images give in small single-threaded batches"""
    import torch
    import torchvision
    from clearml_serving_b200 import formats
    native = gpu_native
    torch.manual_seed(1)
    m = _realistic_bn(torchvision.models.resnet18(weights=None, num_classes=10))
    pm = formats.pack_resnet(m, image_hw=(128, 128))
    rng = np.random.default_rng(3)
    sizes = [1 + (i % 3) for i in range(40)] + [1]    # 41 requests, 80 images
    assert sum(sizes) == 80
    x = rng.standard_normal((80, 3, 128, 128)).astype(np.float32)
    model = native.Model(pm.kind, pm.blob, device=0)
    st = native.Stream(model, 80, 0, 2)
    try:
        reqs, row = [], 0
        for n in sizes:
            reqs.append([x[row:row + n]])
            row += n
        for _ in range(2):                             # twice: the slot's copy bookkeeping is reset between uses
            ev, outs, keep = st.infer_batch(reqs)
            st.wait(ev)
            big = np.concatenate([o[0] for o in outs])
        assert big.shape == (80, 10)
        small = []
        for i in range(0, 80, 2):                      # 2 x 196 KB per batch: the single-threaded path
            ev, outs, keep = st.infer_batch([[x[i:i + 2]]])
            st.wait(ev)
            small.append(outs[0][0].copy())
        assert np.array_equal(big, np.concatenate(small))
        with torch.no_grad():
            ref = m(torch.from_numpy(x[:8])).numpy()
        assert np.abs(big[:8] - ref).max() / np.abs(ref).max() <= REL_TOL
    finally:
        st.destroy()
        model.free()


def test_onnx_file_serves_like_the_eager_module(gpu_native, tmp_path):
    """model.onnx (what the reference hands to Triton's ONNX-Runtime backend, triton_helper.py:169-171) read by
    onnx_reader.py and lowered by model_repo.lower_onnx: logits within 1e-3 of torch-CPU-fp32 of the exported network"""
    import torch
    import torchvision
    from clearml_serving_b200 import model_repo
    from tests.test_model_repo import _export_onnx
    torch.manual_seed(2)
    m = _realistic_bn(torchvision.models.resnet18(weights=None, num_classes=12))
    try:
        proto = _export_onnx(m, torch.zeros(1, 3, 64, 64))
    except Exception as ex:
        pytest.skip("torch's ONNX serialiser is not reachable here: {}".format(ex))
    p = tmp_path / "model.onnx"
    p.write_bytes(proto)
    pm = model_repo.load_model(str(p), framework="onnx")
    x = np.random.default_rng(4).standard_normal((6, 3, 64, 64)).astype(np.float32)
    with torch.no_grad():
        ref = m(torch.from_numpy(x)).numpy()
    got = _run(gpu_native, pm, x, 8)
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert got.shape == ref.shape and err <= REL_TOL, "relative error {:.2e}".format(err)
