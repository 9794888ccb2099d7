"""GPU drop-in test: ps_slm_b200.model.slam_model_asr driven exactly like the reference's
forward()/generate() call sites (Multitask/utils/deepspeed_utils.py:205-208,
Multitask/inference_batch.py:146) with the fake encoder/LLM of tests/fakes.py, against what the
UNMODIFIED reference handed to the LLM for the same inputs (tests/golden/model.npz)."""
import numpy as np
import pytest
import torch

import fakes as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.mark.parametrize("name", list(F.CASES))
def test_model_dispatch_matches_reference(dev, golden, name):
    import ps_slm_b200.model as M
    import ps_slm_b200.projector as P
    g = golden["model"]
    inp = F.build_inputs(name)
    encoder, llm, projector, tok, train_config, model_config = F.build_parts(inp, P.PROJECTORS[inp["proj"]])
    model = M.slam_model_asr(encoder, llm, projector, tok, train_config, model_config,
                             encoder_tokenizer=F.FakeCTCTokenizer())
    wsum = sum(float(p.detach().double().sum()) for p in model.parameters())
    assert abs(wsum - float(g[f"{name}_wsum"])) < 1e-6 * max(1.0, abs(wsum)), "seeded weights differ from the golden run"
    model = model.to(dev)
    batch = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in inp["batch"].items()}
    torch.manual_seed(4321)
    if inp["entry"] == "forward":
        if name == "train_audio":
            # the audio training branch must stay on the fused head: no eager ctc_lo, no [B, T, V] softmax
            def _no_eager(*a, **k):
                raise AssertionError("eager ctc_lo / softmax over the [B, T, 25055] posterior on the training path")
            hook = model.encoder.ctc.ctc_lo.register_forward_pre_hook(lambda m, a: _no_eager())
            real_softmax = torch.softmax
            torch.softmax = lambda x, *a, **k: _no_eager() if (x.dim() == 3 and x.shape[-1] == F.V) else real_softmax(x, *a, **k)
            try:
                out, acc = model(**batch)
            finally:
                torch.softmax = real_softmax
                hook.remove()
        else:
            out, acc = model(**batch)
        assert abs(float(out.loss) - float(g[f"{name}_ref_loss"])) < 2e-2 * max(1.0, abs(float(g[f"{name}_ref_loss"])))
        out.loss.backward()                       # the training call site back-propagates into the projector
        assert all(p.grad is not None for p in model.encoder_projector.parameters())
    else:
        model.generate(**batch)
    seen = llm.seen
    assert np.array_equal(seen["attention_mask"].cpu().numpy(), g[f"{name}_ref_mask"])
    if f"{name}_ref_labels" in g.files:
        assert np.array_equal(seen["labels"].cpu().numpy(), g[f"{name}_ref_labels"])
    if f"{name}_ref_pos" in g.files:
        assert np.array_equal(seen["position_ids"].cpu().numpy(), g[f"{name}_ref_pos"])
    e, r = seen["inputs_embeds"].detach().float().cpu(), torch.from_numpy(g[f"{name}_ref_embeds"])
    assert e.shape == r.shape
    err = ((e - r).norm() / r.norm()).item()
    assert err < 1e-2, f"{name}: inputs_embeds relative error {err}"


def test_raw_feature_psd_from_encoder(dev):
    """bridge.psd_from_encoder (raw-feature branch, no [B,T,V] posterior) vs the oracle's psd on the fp32 posterior of
    the same planted batch: compressed lengths bit-exact, pooled 512-d features to 1e-5."""
    import ps_slm_b200.bridge as bridge
    import ps_slm_b200.synth as S
    from oracle import tasu_oracle as O
    w, b = S.make_ctc_head()
    raw, raw_lens, _ = S.make_encoder_batch(5, 120, w, seed=77, ragged=True)
    raw_lens[3] = 4                                             # an utterance with zero valid frames
    post = torch.softmax(torch.nn.functional.linear(raw, w, b), -1)[:, 4:]
    lens = torch.clamp(raw_lens - 4, min=0)
    ref, ref_lens, _ = O.psd_vec(raw[:, 4:], lens, post, 0, 0.9)
    out, new_lens = bridge.psd_from_encoder(raw.to(dev), raw_lens.to(dev), bridge.cast_weight_bf16(w.to(dev)), b.to(dev))
    assert torch.equal(new_lens.cpu(), ref_lens) and out.shape == ref.shape
    assert torch.allclose(out.cpu(), ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("do_psd,top1", [(True, False), (False, False), (True, True), (False, True)])
def test_voca_trans_project(dev, do_psd, top1):
    """bridge.voca_trans_project (ps-slm.py:485-516 with the evident input) vs the oracle restatement: CTC head over a
    small "LLM vocabulary" (blank = last class), PSD on the logits, softmax(no-blank) · embedding table / top-1 rows."""
    import types

    import ps_slm_b200.bridge as bridge
    import ps_slm_b200.projector as P
    from oracle import tasu_oracle as O
    torch.manual_seed(3)
    Denc, k, Vh, H, B, T = 24, 2, 301, 64, 3, 41                     # head: 300 labels + blank (id 300)
    proj = P.EncoderProjectorLinear(types.SimpleNamespace(encoder_dim=Denc, llm_dim=Vh, encoder_projector_ds_rate=k))
    with torch.no_grad():
        proj.map.weight.mul_(14.0)
        proj.map.bias.zero_()
        proj.map.bias[Vh - 1] = 2.5
    table = torch.randn(Vh + 7, H) * 0.3
    x = torch.randn(B, T, Denc)
    x[:, 1::3] = x[:, 0:-1:3][:, :x[:, 1::3].shape[1]]                # repeated frames → multi-frame runs after the k-concat
    lens = torch.tensor([T, 17, 30])
    with torch.no_grad():
        ref, ref_lens = O.voca_trans(x, lens, proj.map.weight, proj.map.bias, k, table, do_psd, top1, blank_id=Vh - 1)
        out, new_lens = bridge.voca_trans_project(proj.to(dev).eval(), x.to(dev), lens.to(dev), table.to(dev).bfloat16(),
                                                  do_psd, top1, blank_id=Vh - 1)
    assert torch.equal(new_lens.cpu(), ref_lens) and out.shape == ref.shape
    for b in range(B):                                                # rows beyond the compressed length are padding
        n = int(ref_lens[b])
        if n:
            err = ((out[b, :n].float().cpu() - ref[b, :n]).norm() / ref[b, :n].norm()).item()
            assert err < (2e-2 if not top1 else 1e-2), (b, err)


@pytest.mark.parametrize("do_psd", [True, False])
def test_voca_trans_backward(dev, do_psd):
    """Gradient of the voca_trans branch to the CTC head (map.weight / map.bias) vs torch autograd through the oracle:
    dProbs = dOut·Eᵀ, dS = P∘(dProbs − dOut·Out), dW = dSᵀ·x̄ with x̄ the same segmented mean of the input frames."""
    import types

    import ps_slm_b200.bridge as bridge
    import ps_slm_b200.projector as P
    from oracle import tasu_oracle as O
    torch.manual_seed(11)
    Denc, k, Vh, H, B, T = 24, 2, 301, 64, 3, 41
    proj = P.EncoderProjectorLinear(types.SimpleNamespace(encoder_dim=Denc, llm_dim=Vh, encoder_projector_ds_rate=k))
    with torch.no_grad():
        proj.map.weight.mul_(14.0)
        proj.map.bias.zero_()
        proj.map.bias[Vh - 1] = 2.5
    table = torch.randn(Vh + 7, H) * 0.3
    x = torch.randn(B, T, Denc)
    x[:, 1::3] = x[:, 0:-1:3][:, :x[:, 1::3].shape[1]]
    lens = torch.tensor([T, 17, 30])
    w = proj.map.weight.detach().clone().requires_grad_(True)
    b = proj.map.bias.detach().clone().requires_grad_(True)
    ref, ref_lens = O.voca_trans(x, lens, w, b, k, table, do_psd, False, blank_id=Vh - 1)
    valid = (torch.arange(ref.shape[1])[None, :] < ref_lens[:, None]).float().unsqueeze(-1)   # padding rows carry no loss
    gz = torch.randn_like(ref) * valid
    (ref * gz).sum().backward()
    md = proj.to(dev).train()
    out, new_lens = bridge.voca_trans_project(md, x.to(dev), lens.to(dev), table.to(dev).bfloat16(), do_psd, False,
                                              blank_id=Vh - 1)
    assert out.requires_grad and torch.equal(new_lens.cpu(), ref_lens)
    (out * gz.to(dev)).sum().backward()
    for name, g, r in (("map.weight", md.map.weight.grad, w.grad), ("map.bias", md.map.bias.grad, b.grad)):
        assert g is not None and g.shape == r.shape, name
        err = ((g.cpu() - r).norm() / r.norm()).item()
        assert err < 3e-2, f"{name}: relative gradient error {err}"


@pytest.mark.parametrize("do_psd,top1", [(True, False), (False, False), (True, True)])
def test_voca_trans_fused_equals_materialised(dev, do_psd, top1):
    """The voca_trans branch without the logits tensor (head statistics out of the GEMM epilogue, mean-pooling moved to
    the input side of the linear head) against the round-1 formulation that materialises ``[B, T', V]`` logits: same
    lengths, same rows — INCLUDING the rows beyond an utterance's compressed length (uniform mixture / row 0)."""
    import types

    import ps_slm_b200.bridge as bridge
    import ps_slm_b200.projector as P
    torch.manual_seed(5)
    Denc, k, Vh, H, B, T = 64, 2, 1001, 128, 4, 61
    proj = P.EncoderProjectorLinear(types.SimpleNamespace(encoder_dim=Denc, llm_dim=Vh, encoder_projector_ds_rate=k))
    with torch.no_grad():
        proj.map.weight.mul_(10.0)
        proj.map.bias.normal_(0, 0.3)
        proj.map.bias[Vh - 1] = 2.5
    table = (torch.randn(Vh + 5, H) * 0.3).bfloat16().to(dev)
    x = torch.randn(B, T, Denc)
    x[:, 1::3] = x[:, 0:-1:3][:, :x[:, 1::3].shape[1]]
    lens = torch.tensor([T, 23, 40, 9])
    md = proj.to(dev).eval()
    saved = bridge.FUSED_VOCA_TRANS
    try:
        with torch.no_grad():
            bridge.FUSED_VOCA_TRANS = False
            a, la = bridge.voca_trans_project(md, x.to(dev), lens.to(dev), table, do_psd, top1, blank_id=Vh - 1)
            bridge.FUSED_VOCA_TRANS = True
            f, lf = bridge.voca_trans_project(md, x.to(dev), lens.to(dev), table, do_psd, top1, blank_id=Vh - 1)
    finally:
        bridge.FUSED_VOCA_TRANS = saved
    assert torch.equal(la, lf) and a.shape == f.shape
    if top1:
        assert (a.float() - f.float()).abs().max().item() == 0.0 or \
            ((a.float() - f.float()).abs().amax(-1) > 0).float().mean().item() < 0.02    # ties of near-equal logits
    else:
        assert ((a - f).norm() / a.norm()).item() < 1e-2
        pad = torch.arange(a.shape[1], device=dev)[None, :] >= la[:, None]
        if do_psd and bool(pad.any()):
            assert ((a[pad] - f[pad]).abs().max().item()) < 2e-3
