"""GPU parity of the token-row projector (csrc/tokrows.cu): the linear-silu projector applied to
text-simulated posteriors given as (token, hot, base) descriptors, against the fp32 CPU oracle on the
DENSE simulated posterior (ps-slm.py:337-409 → projector.py:149-151) and its autograd gradients."""
import types

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import tasu_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _cfg(D, H):
    return types.SimpleNamespace(encoder_dim=D, llm_dim=H, encoder_projector_ds_rate=1)


def _rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def _ref_module(sd, D, H, Hb=2048):
    norm = nn.LayerNorm(D)
    l1, l2 = nn.Linear(D, Hb), nn.Linear(Hb, H)
    with torch.no_grad():
        norm.weight.copy_(sd["norm.weight"]); norm.bias.copy_(sd["norm.bias"])
        l1.weight.copy_(sd["ffn.0.weight"]); l1.bias.copy_(sd["ffn.0.bias"])
        l2.weight.copy_(sd["ffn.2.weight"]); l2.bias.copy_(sd["ffn.2.bias"])
    return norm, l1, l2


def _module(V, H, seed=0):
    import ps_slm_b200.projector as P
    torch.manual_seed(seed)
    m = P.EncoderProjectorLinearSiLU(_cfg(V, H))
    with torch.no_grad():
        m.norm.weight.uniform_(0.7, 1.3); m.norm.bias.uniform_(-0.1, 0.1); m.ffn[2].bias.uniform_(-0.1, 0.1)
    return m


def _packed(post, lens):
    return torch.cat([post[b, :int(n)] for b, n in enumerate(lens)], 0)


@pytest.mark.parametrize("V,insert_prob", [(300, 0.0), (300, 0.3), (25055, 0.1)])
def test_preactivation_is_fp32_exact(dev, V, insert_prob):
    """z = W1·LN(x) + b1 from the column gather vs the dense fp32 evaluation: 1e-5 (fp32 bar)."""
    import ps_slm_b200.ops as ops
    import ps_slm_b200.sim as sim
    m = _module(V, 96, seed=V)
    g = np.random.default_rng(V)
    ids = [g.integers(0, V, size=int(n)).tolist() for n in (12, 0, 30, 7)]
    ids[2][3] = ids[2][4] = ids[0][1]                       # repeated tokens across and inside utterances
    torch.manual_seed(21)
    post, lens = O.sim_posterior_noise(ids, V, 0, insert_prob=insert_prob)
    torch.manual_seed(21)
    desc = sim.draw_noise_descriptors(ids, V, 0, insert_prob=insert_prob)
    assert desc[3] == lens.tolist()
    with torch.no_grad():
        z_ref = m.ffn[0](m.norm(_packed(post, lens)))
    md = m.to(dev)
    rows = ops.group_token_rows(*desc, V, dev)
    S, D = ops.linear_rowdots(md.ffn[0].weight.detach(), md.norm.weight.detach(), md.norm.bias.detach(), md.ffn[0].bias.detach())
    z, h, _, _ = ops.tokrow_fwd(md.ffn[0].weight.detach(), md.norm.weight.detach(), S, D, rows, md.norm.eps)
    assert z.shape == z_ref.shape
    assert _rel(z.cpu(), z_ref) < 1e-5
    assert _rel(h.float().cpu(), torch.nn.functional.silu(z_ref)) < 4e-3           # bf16 rounding of h only
    # training forward: one dense pass over W1 (compact columns + S, D), then the same row kernel
    colT, S2, D2 = ops.tokrow_cols(md.ffn[0].weight.detach(), md.norm.weight.detach(), md.norm.bias.detach(),
                                   md.ffn[0].bias.detach(), rows)
    assert _rel(S2.cpu(), S.cpu()) < 1e-5 and _rel(D2.cpu(), D.cpu()) < 1e-5
    z2, _, _, _ = ops.tokrow_fwd(md.ffn[0].weight.detach(), md.norm.weight.detach(), S2, D2, rows, md.norm.eps, colT=colT)
    assert _rel(z2.cpu(), z_ref) < 1e-5


def test_clean_rows_inference(dev):
    """generate() branch: clean one-hot simulator (ps-slm.py:337-358) → projector, no grad."""
    import ps_slm_b200.ops as ops
    import ps_slm_b200.sim as sim
    V, H = 25055, 1536
    m = _module(V, H, seed=4)
    ids = [[5, 7, 7, 100, 25054, 3, 3, 9], [1], [], [44, 45, 46, 0, 48]]
    post, lens = O.sim_posterior_clean(ids, V)
    with torch.no_grad():
        y_ref = m(_packed(post, lens)[None])[0] if False else m.ffn(m.norm(_packed(post, lens)))
    md = m.to(dev).eval()
    rows = ops.group_token_rows(*sim.clean_descriptors(ids), V, dev)
    with torch.no_grad():
        y = md.forward_token_rows(rows)
    assert not y.requires_grad and y.shape == y_ref.shape
    assert _rel(y.cpu(), y_ref) < 1e-2
    # zero rows: empty batch
    rows0 = ops.group_token_rows(*sim.clean_descriptors([[], []]), V, dev)
    with torch.no_grad():
        assert md.forward_token_rows(rows0).shape == (0, H)


@pytest.mark.parametrize("V,H,insert_prob", [(300, 96, 0.2), (25055, 1536, 0.0), (25055, 1536, 0.1)])
def test_token_row_backward(dev, V, H, insert_prob):
    """Parameter gradients of the token-row path vs torch autograd through the dense fp32 oracle."""
    import ps_slm_b200.ops as ops
    import ps_slm_b200.sim as sim
    m = _module(V, H, seed=V + 1)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    g = np.random.default_rng(5)
    ids = [g.integers(0, V, size=int(n)).tolist() for n in (25, 3, 40, 0, 18)]
    ids[0][2] = ids[2][0] = ids[2][1] = ids[4][5] = 17       # one token shared by several rows
    torch.manual_seed(33)
    post, lens = O.sim_posterior_noise(ids, V, 0, insert_prob=insert_prob)
    norm, l1, l2 = _ref_module(sd, V, H)
    y_ref = l2(torch.nn.functional.silu(l1(norm(_packed(post, lens)))))
    gy = torch.randn_like(y_ref)
    (y_ref * gy).sum().backward()
    ref = {"norm.weight": norm.weight.grad, "norm.bias": norm.bias.grad, "ffn.0.weight": l1.weight.grad,
           "ffn.0.bias": l1.bias.grad, "ffn.2.weight": l2.weight.grad, "ffn.2.bias": l2.bias.grad}
    md = m.to(dev).train()
    torch.manual_seed(33)
    rows = ops.group_token_rows(*sim.draw_noise_descriptors(ids, V, 0, insert_prob=insert_prob), V, dev)
    y = md.forward_token_rows(rows)
    assert y.requires_grad and _rel(y.detach().cpu(), y_ref.detach()) < 1e-2
    (y * gy.to(dev)).sum().backward()
    for name, p in md.named_parameters():
        assert p.grad is not None and p.grad.shape == p.shape, name
        err = _rel(p.grad.cpu(), ref[name])
        assert err < 2e-2, f"{name}: relative gradient error {err}"
    # deterministic: a second backward gives bit-identical gradients (grouped sums, no row-order races)
    g1 = {n: p.grad.clone() for n, p in md.named_parameters()}
    for p in md.parameters():
        p.grad = None
    y = md.forward_token_rows(rows)
    (y * gy.to(dev)).sum().backward()
    for n, p in md.named_parameters():
        if n in ("ffn.0.weight",):
            assert torch.equal(p.grad, g1[n]), n


@pytest.mark.parametrize("name", ["train_text_noise", "train_text_clean"])
def test_model_text_branch_uses_token_rows(dev, golden, name, monkeypatch):
    """The drop-in model's text-only branch goes through the token-row path (no dense posterior is built) and
    reproduces what the unmodified reference handed to the LLM (tests/golden/model.npz)."""
    import fakes as F
    import ps_slm_b200.model as M
    import ps_slm_b200.projector as P
    import ps_slm_b200.sim as sim
    g = golden["model"]
    inp = F.build_inputs(name)
    encoder, llm, projector, tok, train_config, model_config = F.build_parts(inp, P.PROJECTORS[inp["proj"]])
    model = M.slam_model_asr(encoder, llm, projector, tok, train_config, model_config,
                             encoder_tokenizer=F.FakeCTCTokenizer()).to(dev)

    def boom(*a, **k):
        raise AssertionError("dense simulated posterior built on the token-row path")
    monkeypatch.setattr(sim, "build_dense", boom)
    batch = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in inp["batch"].items()}
    torch.manual_seed(4321)
    out, _ = model(**batch)
    out.loss.backward()
    assert all(p.grad is not None for p in model.encoder_projector.parameters())
    seen = llm.seen
    assert np.array_equal(seen["attention_mask"].cpu().numpy(), g[f"{name}_ref_mask"])
    assert np.array_equal(seen["labels"].cpu().numpy(), g[f"{name}_ref_labels"])
    e, r = seen["inputs_embeds"].detach().float().cpu(), torch.from_numpy(g[f"{name}_ref_embeds"])
    assert e.shape == r.shape and _rel(e, r) < 1e-2


def test_zero_rows_training_step(dev):
    """Every transcript dropped / empty: forward gives [0, H], backward gives all-zero parameter gradients."""
    import ps_slm_b200.ops as ops
    import ps_slm_b200.sim as sim
    V, H = 300, 96
    m = _module(V, H).to(dev).train()
    rows = ops.group_token_rows(*sim.clean_descriptors([[], [], []]), V, dev)
    y = m.forward_token_rows(rows)
    assert y.shape == (0, H) and y.requires_grad
    y.sum().backward()
    for n, p in m.named_parameters():
        assert p.grad is not None and float(p.grad.abs().sum()) == 0.0, n
