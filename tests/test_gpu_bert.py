"""GPU parity (-m gpu) for the DL graph path (BASELINE.json configs[3]: BERT-base fp16, mixed sequence
lengths, ragged batching) against the torch-CPU-fp32 forward of the SAME weights -- the stand-in for
the reference's Triton-CPU libtorch backend (parity unpinned by reference tests).
Tolerance: max |logit - ref| / max |ref| <= 1e-3 (north_star: "within 1e-3 rel for fp32 DL models")."""
import asyncio

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REL_TOL = 1e-3


def _make(cfg_kwargs, seed=0):
    import torch
    from transformers import BertConfig, BertForSequenceClassification
    torch.manual_seed(seed)
    model = BertForSequenceClassification(BertConfig(**cfg_kwargs)).eval()
    # random-init classifier/pooler weights are tiny (std 0.02): scale them so that the logits have O(1)
    # magnitude and the relative tolerance is meaningful
    with torch.no_grad():
        model.classifier.weight.mul_(20.0)
        model.classifier.bias.normal_(0, 0.5)
    return model


def _requests(lens, vocab, seed=1):
    rng = np.random.default_rng(seed)
    reqs = []
    for n in lens:
        ids = rng.integers(0, vocab, (1, n)).astype(np.int32)
        tt = np.zeros((1, n), np.int32)
        tt[0, n // 2:] = 1
        mask = np.ones((1, n), np.int32)
        reqs.append([ids, tt, mask])
    return reqs


def _reference(model, reqs):
    import torch
    out = []
    with torch.no_grad():
        for ids, tt, mask in reqs:
            r = model(input_ids=torch.from_numpy(ids).long(), token_type_ids=torch.from_numpy(tt).long(),
                      attention_mask=torch.from_numpy(mask).long())
            out.append(r.logits.float().numpy())
    return np.concatenate(out)


def _run(native, model_t, reqs, max_rows=64, max_seq=256):
    from clearml_serving_b200 import formats
    pm = formats.pack_bert(model_t)
    model = native.Model(pm.kind, pm.blob, device=0)
    st = native.Stream(model, max_rows, max_seq, 2)
    try:
        ev, outs, keep = st.infer_batch(reqs)
        st.wait(ev)
        got = np.concatenate([o[0] for o in outs])
        # a second, differently composed batch must give the same rows (no dependence on batch-mates)
        ev, outs2, keep = st.infer_batch(reqs[::-1][:3])
        st.wait(ev)
        for k, o in enumerate(outs2):
            idx = len(reqs) - 1 - k
            assert np.array_equal(o[0], got[idx:idx + 1]), "result depends on batch composition"
        return got
    finally:
        st.destroy()
        model.free()


def test_small_bert_parity(gpu_native):
    model_t = _make(dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=512,
                         vocab_size=1000, max_position_embeddings=256))
    reqs = _requests([5, 64, 17, 128, 1, 200, 33], 1000)
    ref = _reference(model_t, reqs)
    got = _run(gpu_native, model_t, reqs)
    assert got.shape == ref.shape and got.dtype == np.float32
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err <= REL_TOL, "relative error {:.2e}".format(err)


def test_bert_base_parity_mixed_lengths(gpu_native):
    """configs[3]: BERT-base (12L, H768, 12 heads, FFN 3072, vocab 30522, 2 labels), S in {16,64,128,256}."""
    model_t = _make(dict())
    reqs = _requests([16, 256, 64, 128, 16, 64, 256, 128], 30522)
    ref = _reference(model_t, reqs)
    got = _run(gpu_native, model_t, reqs)
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err <= REL_TOL, "relative error {:.2e}".format(err)


def test_bert_base_parity_at_the_baseline_batch(gpu_native):
    """BASELINE.json configs[3] at its FULL batch: 64 requests, S drawn from {16, 64, 128, 256} (the bench's mix), every
    row against the torch-CPU-fp32 forward of the same weights -- and the same rows again inside a different batch."""
    import torch
    torch.set_num_threads(max(1, min(32, torch.get_num_threads())))
    model_t = _make(dict())
    lens = np.random.default_rng(1).choice([16, 64, 128, 256], size=64).tolist()
    reqs = _requests(lens, 30522, seed=11)
    ref = _reference(model_t, reqs)
    got = _run(gpu_native, model_t, reqs)
    assert got.shape == (64, 2)
    err = np.abs(got - ref).max(axis=1) / np.abs(ref).max()
    assert err.max() <= REL_TOL, "worst row {} (S={}): relative error {:.2e}".format(int(err.argmax()), lens[int(err.argmax())], err.max())


def test_bert_masked_keys_and_engine_api(gpu_native, tmp_path):
    """attention_mask with padding inside a request + the plugin API (HF example's preprocess output:
    three lists, examples/huggingface/preprocess.py:23)."""
    import torch
    from clearml_serving_b200 import BasePreprocessRequest, ModelEndpoint
    model_t = _make(dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=512,
                         vocab_size=500, max_position_embeddings=64))
    model_t.save_pretrained(tmp_path / "bert")
    ep = ModelEndpoint(engine_type="b200", serving_url="transformer_model", model_id=str(tmp_path / "bert"),
                       input_size=[[-1], [-1], [-1]], input_type=["int32", "int32", "int32"],
                       input_name=["input_ids", "token_type_ids", "attention_mask"],
                       output_size=[[2]], output_type=["float32"], output_name=["output"],
                       auxiliary_cfg={"max_batch_size": 16, "dynamic_batching.max_queue_delay_microseconds": 2000,
                                      "b200.max_seq_len": 64})
    eng = BasePreprocessRequest.get_engine_cls("b200")(model_endpoint=ep, task=None)
    rng = np.random.default_rng(5)
    bodies, refs = [], []
    for n in (12, 40, 7, 64, 23):
        ids = rng.integers(0, 500, (1, n))
        mask = np.ones((1, n), np.int64)
        mask[0, n - n // 4:] = 0          # right padding, as a tokenizer with padding=True would emit
        tt = np.zeros((1, n), np.int64)
        bodies.append([ids.tolist(), tt.tolist(), mask.tolist()])
        with torch.no_grad():
            refs.append(model_t(input_ids=torch.from_numpy(ids), token_type_ids=torch.from_numpy(tt),
                                attention_mask=torch.from_numpy(mask)).logits.numpy())

    async def main():
        return await asyncio.gather(*[eng.process(b, {}, None) for b in bodies])
    try:
        outs = asyncio.run(main())
        for o, r in zip(outs, refs):
            assert o.shape == (1, 2) and o.dtype == np.float32
            assert np.abs(o - r).max() / np.abs(np.concatenate(refs)).max() <= REL_TOL
        with pytest.raises(ValueError, match="tokens is outside"):
            eng.process_sync([[list(range(65))], [[0] * 65], [[1] * 65]])
    finally:
        eng.unload()
