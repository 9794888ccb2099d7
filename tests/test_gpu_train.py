"""GPU parity tests of the training path (projector backward, splice backward, text-only step)
against torch autograd through the fp32 oracle on the CPU."""
import types

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import tasu_oracle as O

pytestmark = pytest.mark.gpu
SP, PAD = 151665, 151643


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _cfg(D, H, k=1):
    return types.SimpleNamespace(encoder_dim=D, llm_dim=H, encoder_projector_ds_rate=k)


def _ref_module(sd, D, H):
    norm = nn.LayerNorm(D)
    l1, l2 = nn.Linear(D, 2048), nn.Linear(2048, H)
    with torch.no_grad():
        norm.weight.copy_(sd["norm.weight"]); norm.bias.copy_(sd["norm.bias"])
        l1.weight.copy_(sd["ffn.0.weight"]); l1.bias.copy_(sd["ffn.0.bias"])
        l2.weight.copy_(sd["ffn.2.weight"]); l2.bias.copy_(sd["ffn.2.bias"])
    return norm, l1, l2


def _rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


@pytest.mark.parametrize("D,H,B,T", [(300, 96, 2, 30), (25055, 1536, 2, 90)])
def test_linear_silu_backward(dev, D, H, B, T):
    import ps_slm_b200.projector as P
    torch.manual_seed(D)
    m = P.EncoderProjectorLinearSiLU(_cfg(D, H))
    with torch.no_grad():
        m.norm.weight.uniform_(0.7, 1.3); m.norm.bias.uniform_(-0.1, 0.1); m.ffn[2].bias.uniform_(-0.1, 0.1)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    x = torch.softmax(torch.randn(B, T, D) * 5, -1)
    x[1, T - 5:] = 0
    gy = torch.randn(B, T, H)
    norm, l1, l2 = _ref_module(sd, D, H)
    y_ref = l2(torch.nn.functional.silu(l1(norm(x))))
    (y_ref * gy).sum().backward()
    ref = {"norm.weight": norm.weight.grad, "norm.bias": norm.bias.grad, "ffn.0.weight": l1.weight.grad,
           "ffn.0.bias": l1.bias.grad, "ffn.2.weight": l2.weight.grad, "ffn.2.bias": l2.bias.grad}
    m = m.to(dev).train()
    y = m(x.to(dev))
    assert y.requires_grad and _rel(y.detach().cpu(), y_ref.detach()) < 1e-2
    (y * gy.to(dev)).sum().backward()
    for name, p in m.named_parameters():
        assert p.grad is not None and p.grad.shape == p.shape, name
        err = _rel(p.grad.cpu(), ref[name])
        assert err < 2e-2, f"{name}: relative gradient error {err}"
    # a second step sees the updated parameters (cache keyed on the version counter)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(0.01 * torch.randn_like(p))
    sd2 = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    norm, l1, l2 = _ref_module(sd2, D, H)
    y2 = m(x.to(dev)).detach().cpu()
    assert _rel(y2, l2(torch.nn.functional.silu(l1(norm(x)))).detach()) < 1e-2


@pytest.mark.parametrize("kind", ["linear", "simple_linear"])
def test_other_projectors_backward(dev, kind):
    import ps_slm_b200.projector as P
    torch.manual_seed(3)
    D, H, k, B, T = 40, 72, 2, 3, 21
    cls = P.EncoderProjectorConcat if kind == "linear" else P.EncoderProjectorLinear
    m = cls(_cfg(D, H, k))
    with torch.no_grad():           # bf16-representable operands: the ReLU mask (a discontinuity) is then identical
        for p in m.parameters():    # in both paths and the comparison measures the contractions, not mask flips
            p.copy_(p.bfloat16().float())
    sd = {n: v.clone() for n, v in m.state_dict().items()}
    x = torch.randn(B, T, D).bfloat16().float().requires_grad_(True)
    if kind == "linear":
        y_ref = O.projector_concat(x, k, sd["linear1.weight"].requires_grad_(), sd["linear1.bias"].requires_grad_(),
                                   sd["linear2.weight"].requires_grad_(), sd["linear2.bias"].requires_grad_())
    else:
        y_ref = O.projector_linear(x, k, sd["map.weight"].requires_grad_(), sd["map.bias"].requires_grad_())
    gy = torch.randn_like(y_ref)
    (y_ref * gy).sum().backward()
    m = m.to(dev).train()
    xd = x.detach().to(dev).requires_grad_(True)
    y = m(xd)
    assert y.shape == y_ref.shape and _rel(y.detach().cpu(), y_ref.detach()) < 1e-2
    (y * gy.to(dev)).sum().backward()
    # two chained bf16 contractions with a tiny K (80): quantisation noise is relatively larger than at 25055
    for name, p in m.named_parameters():
        err = _rel(p.grad.cpu(), sd[name].grad)
        assert err < 3e-2, f"{name}: {err}"
    assert _rel(xd.grad.cpu(), x.grad) < 3e-2


def test_splice_backward(dev):
    import ps_slm_b200.bridge as bridge
    torch.manual_seed(2)
    B, S, H = 3, 12, 64
    ids = torch.randint(1, 5000, (B, S)); att = torch.ones(B, S, dtype=torch.bool)
    ids[0, 3] = SP; ids[1, 0] = SP; ids[2, 7] = SP
    att[1, 9:] = False; ids[1, 9:] = PAD; att[2, 11:] = False; ids[2, 11:] = PAD
    M = torch.tensor([4, 0, 7])
    emb = torch.randn(B, S, H)
    af = torch.randn(B, 7, H, requires_grad=True)
    lab = torch.randint(0, 100, (B, S))
    e_r, _, _, _, _ = O.merge(af, M, emb, ids, att, lab, SP, PAD)
    R = torch.randn_like(e_r)
    (e_r * R).sum().backward()
    afd = af.detach().to(dev).requires_grad_(True)
    out = bridge.merge_input_ids_with_audio_features(afd, M.to(dev), emb.to(dev), ids.to(dev), att.to(dev), lab.to(dev), SP, PAD)
    assert torch.equal(out[0].detach().cpu(), e_r.detach())
    (out[0] * R.to(dev)).sum().backward()
    assert torch.equal(afd.grad.cpu(), af.grad)
    # the text embeddings are trained too (freeze_llm=False / PEFT with hot embed_tokens, ps-slm.py:119-123): the
    # index_put of ps-slm.py:833 hands every text token the gradient of the row it was copied to
    emb_r = emb.clone().requires_grad_(True)
    af_r = af.detach().clone().requires_grad_(True)
    e_r2 = O.merge(af_r, M, emb_r, ids, att, lab, SP, PAD)[0]
    (e_r2 * R).sum().backward()
    embd = emb.to(dev).clone().requires_grad_(True)
    afd2 = af.detach().to(dev).clone().requires_grad_(True)
    out2 = bridge.merge_input_ids_with_audio_features(afd2, M.to(dev), embd, ids.to(dev), att.to(dev), lab.to(dev), SP, PAD)
    (out2[0] * R.to(dev)).sum().backward()
    assert torch.equal(embd.grad.cpu(), emb_r.grad) and torch.equal(afd2.grad.cpu(), af_r.grad)
    # only the text embeddings require a gradient (frozen projector)
    embd3 = emb.to(dev).clone().requires_grad_(True)
    out3 = bridge.merge_input_ids_with_audio_features(af.detach().to(dev), M.to(dev), embd3, ids.to(dev), att.to(dev), lab.to(dev), SP, PAD)
    (out3[0] * R.to(dev)).sum().backward()
    assert torch.equal(embd3.grad.cpu(), emb_r.grad)
    # fused embedding lookup with a trainable table: the lookup falls back to the differentiable torch op
    table = (torch.randn(SP + 1, emb.shape[-1]) * 0.1)          # embed_tokens also looks the <speech> id up (ps-slm.py:525)
    ids_small = ids
    t_r = table.clone().requires_grad_(True)
    e_r4 = O.merge(af.detach(), M, torch.nn.functional.embedding(ids_small, t_r), ids_small, att, lab, SP, PAD)[0]
    (e_r4 * R).sum().backward()
    t_d = table.to(dev).clone().requires_grad_(True)
    packed = torch.cat([af.detach()[b, :int(M[b])] for b in range(af.shape[0])]).to(dev)
    out4 = bridge.merge_packed_audio_rows(packed, M.to(dev), int(M.max()), t_d, 1, ids_small.to(dev), att.to(dev), lab.to(dev), SP, PAD)
    (out4[0] * R.to(dev)).sum().backward()
    assert torch.allclose(t_d.grad.cpu(), t_r.grad, rtol=1e-5, atol=1e-6)


def test_text_only_training_step(dev):
    """BASELINE config 3 in miniature: simulated posteriors (reference RNG order) → projector fwd/bwd
    → splice with labels, loss = <inputs_embeds, R>; parameter gradients vs the fp32 CPU oracle."""
    import ps_slm_b200.projector as P
    import ps_slm_b200.sim as sim
    import ps_slm_b200.synth as S
    import ps_slm_b200.ops as ops
    from ps_slm_b200.autograd import SpliceFunction, linear_silu_train_rows
    B, V, H = 4, S.V_CTC, S.H_LLM
    ids_list = S.make_transcripts(B, V, seed=5, lo=8, hi=20)
    input_ids, mask, labels = S.make_prompts(B, seed=9, left_pad=False, target_lens=[len(i) for i in ids_list])
    torch.manual_seed(0)
    m = P.EncoderProjectorLinearSiLU(_cfg(V, H))
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    table = (torch.randn(S.V_LLM, H) * 0.02)
    # ---- CPU oracle
    torch.manual_seed(77)
    post, lens = O.sim_posterior_noise(ids_list, V, 0, insert_prob=0.1)
    norm, l1, l2 = _ref_module(sd, V, H)
    proj = l2(torch.nn.functional.silu(l1(norm(post))))
    e_r, m_r, l_r, p_r, _ = O.merge(proj, lens, torch.nn.functional.embedding(input_ids, table), input_ids, mask,
                                    labels, S.SPEECH_ID, S.PAD_ID)
    R = torch.randn_like(e_r)
    (e_r * R).sum().backward()
    ref = {"norm.weight": norm.weight.grad, "norm.bias": norm.bias.grad, "ffn.0.weight": l1.weight.grad,
           "ffn.0.bias": l1.bias.grad, "ffn.2.weight": l2.weight.grad, "ffn.2.bias": l2.bias.grad}
    # ---- B200 path: packed bf16 rows straight from the decisions
    m = m.to(dev).train()
    torch.manual_seed(77)
    dec = sim.draw_noise_decisions(ids_list, 0, insert_prob=0.1)
    rows, mean, rstd, lens_d = sim.build_packed_bf16(dec, V, dev)
    assert torch.equal(lens_d.cpu(), lens)
    y = linear_silu_train_rows(m, rows, mean, rstd, rows.shape[0], torch.float32)
    sp = ops.splice_rowstat(input_ids.to(dev), mask.to(dev), S.SPEECH_ID)
    ops.splice_plan(sp, lens_d, 1)
    hdr = sp.header.cpu()
    emb, mk, lb, pos, _ = SpliceFunction.apply(y, table.to(dev), sp, int(hdr[0]), 1, 0, int(lens.max()),
                                               labels.to(dev), S.PAD_ID, S.IGNORE_ID)
    assert torch.equal(mk.cpu(), m_r) and torch.equal(lb.cpu(), l_r) and torch.equal(pos.cpu(), p_r)
    assert _rel(emb.detach().cpu(), e_r.detach()) < 1e-2
    (emb * R.to(dev)).sum().backward()
    for name, p in m.named_parameters():
        err = _rel(p.grad.cpu(), ref[name])
        assert err < 2e-2, f"{name}: relative gradient error {err}"
