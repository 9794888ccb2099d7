"""GPU parity at BASELINE.json's full sizes (configs[1]: B=64 x 30 s) through size-independent
properties, plus a mid-size run against the vectorised oracle."""
import types

import numpy as np
import pytest
import torch

from oracle import tasu_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _bridge(dev, table_dtype=torch.bfloat16):
    import ps_slm_b200.projector as P
    import ps_slm_b200.synth as S
    from ps_slm_b200.bridge import TasuBridge
    w, b = S.make_ctc_head()
    torch.manual_seed(0)
    cfg = types.SimpleNamespace(encoder_dim=S.V_CTC, llm_dim=S.H_LLM, encoder_projector_ds_rate=1)
    proj = P.EncoderProjectorLinearSiLU(cfg).to(dev).eval()
    table = S.make_embed_table(dtype=table_dtype, device=dev)
    return w, b, proj, table, TasuBridge(w.to(dev), b.to(dev), proj, table, S.SPEECH_ID, S.PAD_ID)


@pytest.mark.parametrize("T,ragged", [(500, False), (1000, True), (83, True)])
def test_full_batch_properties(dev, T, ragged):
    """B=64: greedy ids == planted labels, kept candidates == first-principles plan, splice invariants,
    fused and materialised paths agree on every integer and within 1e-2 on the embeddings."""
    import ps_slm_b200.ops as ops
    import ps_slm_b200.synth as S
    import ps_slm_b200._lib as L
    B = 64
    w, b, proj, table, br = _bridge(dev)
    raw, raw_lens, lab, soft = S.make_encoder_batch(B, T, w, seed=T, ragged=ragged, return_soft=True)
    ids, mask, _ = S.make_prompts(B, seed=T, left_pad=True)
    lens = raw_lens - 4
    exp = S.expected_plan(lab, soft, lens)
    exp_lens = torch.tensor([len(e) for e in exp])
    outs = {}
    for mode in (False, True):
        br.materialize_logits = mode
        e, m, _, p, nl = br(raw.to(dev), raw_lens.to(dev), ids.to(dev), mask.to(dev))
        outs[mode] = (e, m, p, nl)
        assert torch.equal(nl.cpu(), exp_lens), "compressed lengths differ from the planted plan"
        # splice invariants: one row per prompt token + M_b - 1, left padded, positions count the mask
        Sp = e.shape[1]
        tok = mask.sum(1)
        assert Sp == ids.shape[1] + int(exp_lens.max()) - 1          # pad positions count one each (ps-slm.py:805-809)
        assert torch.equal(m.sum(1).cpu(), tok + exp_lens - 1)
        mm = m.cpu()
        assert bool((mm[:, 1:] >= mm[:, :-1]).all()), "left padding: mask must be non-decreasing"
        pos = p.cpu()
        ref_pos = torch.cumsum(mm.long(), 1) - 1
        ref_pos[~mm] = 1
        assert torch.equal(pos, ref_pos)
        assert bool((e.cpu()[~mm] == 0).all())
    assert torch.equal(outs[False][1], outs[True][1]) and torch.equal(outs[False][2], outs[True][2])
    a, c = outs[False][0].float(), outs[True][0].float()
    assert ((a - c).norm() / c.norm()).item() < 1e-2
    # the integer plan itself
    x2, _, _ = ops.cast_rows(raw.to(dev).reshape(B * (T + 4), 512), torch.bfloat16)
    wq, _, _ = ops.cast_rows(w.to(dev), torch.bfloat16)
    st = ops.ctc_head_stats(x2, wq, b.to(dev), B, T, 4, S.V_CTC, 512, 0)
    am = st.argmax.cpu().view(B, T).long()
    valid = torch.arange(T)[None] < lens[:, None]
    assert torch.equal(am[valid], lab[valid]), "greedy ids differ from the planted labels"
    plan = ops.collapse_plan(st, lens.to(dev), 0, 0.9)
    ss, sl = plan.seg_start.cpu().view(B, T), plan.seg_len.cpu().view(B, T)
    for bb in range(B):
        n = len(exp[bb])
        assert list(zip(ss[bb, :n].tolist(), sl[bb, :n].tolist())) == exp[bb]
    hdr = plan.header.cpu()
    assert int(hdr[L.CH_N_OUT]) == int(exp_lens.sum()) and int(hdr[L.CH_MAX_LEN]) == int(exp_lens.max())
    assert int(hdr[L.CH_KEPT_FRAMES]) == sum(n for e_ in exp for _, n in e_)


def test_midsize_vs_oracle(dev):
    """B=12 x 30 s through the whole bridge against the vectorised fp32 oracle."""
    import ps_slm_b200.synth as S
    B, T = 12, 500
    w, b, proj, table, br = _bridge(dev, torch.float32)
    raw, raw_lens, _ = S.make_encoder_batch(B, T, w, seed=99, ragged=True)
    ids, mask, _ = S.make_prompts(B, seed=99, left_pad=True)
    sd = {k: v.detach().cpu() for k, v in proj.state_dict().items()}
    pp = (sd["norm.weight"], sd["norm.bias"], sd["ffn.0.weight"], sd["ffn.0.bias"], sd["ffn.2.weight"], sd["ffn.2.bias"])
    (e_r, m_r, _, p_r, f_r), nl_r = O.bridge_inference(raw, raw_lens, w, b, pp, table.cpu(), ids, mask, None,
                                                      S.SPEECH_ID, S.PAD_ID)
    e, m, _, p, nl = br(raw.to(dev), raw_lens.to(dev), ids.to(dev), mask.to(dev))
    assert torch.equal(nl.cpu(), nl_r) and torch.equal(m.cpu(), m_r) and torch.equal(p.cpu(), p_r)
    audio = m_r & (f_r == S.PAD_ID)
    e = e.cpu()
    rows = (e[audio] - e_r[audio]).norm(dim=-1) / e_r[audio].norm(dim=-1)
    assert rows.max().item() < 2e-2 and ((e[audio] - e_r[audio]).norm() / e_r[audio].norm()).item() < 1e-2
    assert torch.equal(e[~audio], e_r[~audio])


def test_fullsize_vs_oracle(dev):
    """BASELINE.json configs[1] at its full size (B = 64 x 30 s, ragged lengths) through the whole bridge — the
    benchmarked configuration, exact decisions on — against the ORACLE (oracle/tasu_oracle.py: fp32 softmax(ctc_lo) of the
    full [64, 504, 25055] posterior on the host, psd_vec, fp32 projector, merge): compressed lengths, masks, position
    ids and the text rows bit-exact; projected audio rows within 1e-2 (bf16 tensor-core path)."""
    import ps_slm_b200.synth as S
    B, T = 64, 500
    w, b, proj, table, br = _bridge(dev, torch.float32)
    assert br.exact_decisions
    raw, raw_lens, _ = S.make_encoder_batch(B, T, w, seed=2025, ragged=True)
    ids, mask, _ = S.make_prompts(B, seed=2025, left_pad=True)
    sd = {k: v.detach().cpu() for k, v in proj.state_dict().items()}
    pp = (sd["norm.weight"], sd["norm.bias"], sd["ffn.0.weight"], sd["ffn.0.bias"], sd["ffn.2.weight"], sd["ffn.2.bias"])
    with torch.no_grad():
        post, lens = O.ctc_head_posterior(raw, raw_lens, w, b)
        feats, nl_r, plan = O.psd_vec(post, lens, post, 0)
        del post
        proj_r = O.projector_linear_silu(feats, *pp)
        del feats
        e_r, m_r, _, p_r, f_r = O.merge(proj_r, nl_r, torch.nn.functional.embedding(ids, table.cpu()), ids, mask, None,
                                        S.SPEECH_ID, S.PAD_ID)
    e, m, _, p, nl = br(raw.to(dev), raw_lens.to(dev), ids.to(dev), mask.to(dev))
    assert torch.equal(nl.cpu(), nl_r), "compressed lengths differ from the oracle"
    assert torch.equal(m.cpu(), m_r) and torch.equal(p.cpu(), p_r)
    audio = m_r & (f_r == S.PAD_ID)
    e = e.cpu()
    assert torch.equal(e[~audio], e_r[~audio])
    assert ((e[audio] - e_r[audio]).norm() / e_r[audio].norm()).item() < 1e-2
    rows = (e[audio] - e_r[audio]).norm(dim=-1) / e_r[audio].norm(dim=-1)
    assert rows.max().item() < 3e-2
    # the integer plan of the device (greedy ids of every valid frame, kept candidates) against the oracle's
    st_ids = plan["ids"]
    import ps_slm_b200.ops as ops
    x2, _, _ = ops.cast_rows(raw.to(dev).reshape(B * (T + 4), 512), torch.bfloat16)
    wq, _, _ = ops.cast_rows(w.to(dev), torch.bfloat16)
    st = ops.ctc_head_stats(x2, wq, b.to(dev), B, T, 4, S.V_CTC, 512, 0)
    ops.refine_ambiguous_frames(st, lens.to(dev), raw.to(dev).reshape(B * (T + 4), 512), w.to(dev), b.to(dev),
                                ops.row_norm_max(w.to(dev)), T, 4, S.V_CTC, 0, 0.9)
    valid = torch.arange(T)[None] < lens[:, None]
    assert torch.equal(st.argmax.cpu().view(B, T).long()[valid], st_ids[valid])
    dplan = ops.collapse_plan(st, lens.to(dev), 0, 0.9)
    ss, sl = dplan.seg_start.cpu().view(B, T), dplan.seg_len.cpu().view(B, T)
    kb, kt, kl = plan["kept_b"], plan["kept_t"], plan["kept_len"]
    for bb in range(B):
        sel = kb == bb
        n = int(sel.sum())
        assert ss[bb, :n].tolist() == kt[sel].tolist() and sl[bb, :n].tolist() == kl[sel].tolist()


def test_capacity_overflow_is_redone_exactly(dev):
    """The speculative tail is sized from the high-water mark of earlier batches.  A batch that keeps MORE frames than that
    (a quiet batch followed by a dense one) must neither touch memory beyond the buffers (tasu_pool_tail is bounded by the
    compact matrix's row capacity) nor return a truncated result: the host sees the overflow in the header and redoes the
    tail with exact sizes."""
    import ps_slm_b200.synth as S
    B, T = 16, 500
    w, b, proj, table, br = _bridge(dev)
    raw, raw_lens, _ = S.make_encoder_batch(B, T, w, seed=5)
    ids, mask, _ = S.make_prompts(B, seed=5, left_pad=True)
    args = (raw.to(dev), raw_lens.to(dev), ids.to(dev), mask.to(dev))
    e0, m0, _, p0, nl0 = br(*args)                              # first call: worst-case capacity
    kept = br.last_counts["kept_frames"]
    assert kept > 2048 and br.last_counts["n_out"] > 2048, "the batch must exceed the smallest capacity quantum"
    br._capacity[(B, T)] = (1, 1)                               # pretend every earlier batch was almost empty
    e1, m1, _, p1, nl1 = br(*args)
    torch.cuda.synchronize()
    assert torch.equal(nl0, nl1) and torch.equal(m0, m1) and torch.equal(p0, p1)
    assert torch.equal(e0, e1), "the redone tail must equal the exact-capacity result bit for bit"
    assert br._capacity[(B, T)] == (kept, br.last_counts["n_out"])


def _small_bridge(dev, V=61, D=32, H=48, table_rows=300):
    import ps_slm_b200.projector as P
    from ps_slm_b200.bridge import TasuBridge
    import math
    g = torch.Generator().manual_seed(1)
    w = torch.randn(V, D, generator=g) / math.sqrt(D)
    b = torch.randn(V, generator=g) * 0.05
    torch.manual_seed(2)
    proj = P.EncoderProjectorLinearSiLU(types.SimpleNamespace(encoder_dim=V, llm_dim=H, encoder_projector_ds_rate=1)).to(dev).eval()
    table = torch.randn(table_rows, H, generator=g)
    return w, b, proj, table, TasuBridge(w.to(dev), b.to(dev), proj, table.to(dev), 299, 0)


def _oracle_small(raw, raw_lens, w, b, proj, table, ids, mask, labels=None):
    sd = {k: v.detach().cpu() for k, v in proj.state_dict().items()}
    pp = (sd["norm.weight"], sd["norm.bias"], sd["ffn.0.weight"], sd["ffn.0.bias"], sd["ffn.2.weight"], sd["ffn.2.bias"])
    return O.bridge_inference(raw, raw_lens, w, b, pp, table, ids, mask, labels, 299, 0)


@pytest.mark.parametrize("case", ["all_dropped", "single_frame", "zero_length_rows", "one_utt", "everything_kept"])
def test_bridge_edge_cases(dev, case):
    """Empty / ragged / degenerate inputs through the fused path vs the oracle (ps-slm.py:261-264, :304-306)."""
    import math
    w, b, proj, table, br = _small_bridge(dev)
    V, D = w.shape
    what = w / w.norm(dim=1, keepdim=True)
    g = torch.Generator().manual_seed(7)

    def frames(labels, scale):
        lab = torch.tensor(labels)
        return (scale / w.norm(dim=1)[lab]).unsqueeze(-1) * what[lab] + 0.02 * torch.randn(len(labels), D, generator=g)

    if case == "all_dropped":            # confident blanks only → every M_b = 0, the <speech> slot vanishes
        B, T = 2, 6
        body = torch.stack([frames([0] * T, 16.0) for _ in range(B)])
        lens = torch.tensor([T, T])
    elif case == "single_frame":
        B, T = 1, 1
        body = frames([5], 14.0).unsqueeze(0)
        lens = torch.tensor([1])
    elif case == "zero_length_rows":     # L = 0 for some utterances
        B, T = 3, 9
        body = torch.stack([frames([3, 3, 0, 7, 0, 0, 9, 9, 9], 14.0) for _ in range(B)])
        lens = torch.tensor([9, 0, 4])
    elif case == "one_utt":
        B, T = 1, 12
        body = frames([0, 4, 4, 0, 0, 8, 0, 2, 2, 2, 0, 1], 14.0).unsqueeze(0)
        lens = torch.tensor([12])
    else:                                # all different tokens, nothing merges or drops
        B, T = 2, 10
        body = torch.stack([frames(list(range(1 + i, 11 + i)), 14.0) for i in range(B)])
        lens = torch.tensor([10, 7])
    raw = torch.cat([torch.randn(B, 4, D, generator=g) * 0.1, body], 1).contiguous()
    raw_lens = lens + 4
    ids = torch.randint(1, 298, (B, 6), generator=g)
    ids[:, 2] = 299
    mask = torch.ones(B, 6, dtype=torch.bool)
    if B > 1:
        mask[1, :2] = False
        ids[1, :2] = 0
    (e_r, m_r, _, p_r, f_r), nl_r = _oracle_small(raw, raw_lens, w, b, proj, table, ids, mask)
    e, m, _, p, nl = br(raw.to(dev), raw_lens.to(dev), ids.to(dev), mask.to(dev))
    assert torch.equal(nl.cpu(), nl_r), (nl.cpu(), nl_r)
    assert e.shape == e_r.shape and torch.equal(m.cpu(), m_r) and torch.equal(p.cpu(), p_r)
    if case == "all_dropped":
        assert int(nl_r.sum()) == 0 and e.shape[1] == 5
    if e_r.numel():
        assert ((e.cpu() - e_r).norm() / e_r.norm()).item() < 1e-2
    # the stand-alone psd() on the same posterior agrees too
    import ps_slm_b200.bridge as bridge
    post, plens = O.ctc_head_posterior(raw, raw_lens, w, b)
    f_ref, l_ref = O.psd_loop(post, plens, post)
    f_gpu, l_gpu = bridge.psd(post.to(dev), plens.to(dev), post.to(dev))
    assert torch.equal(l_gpu.cpu(), l_ref) and f_gpu.shape == f_ref.shape
    if f_ref.numel():
        np.testing.assert_allclose(f_gpu.cpu().numpy(), f_ref.numpy(), rtol=1e-5, atol=1e-7)


def test_cached_weight_copies_follow_data_copy(dev):
    """DeepSpeed ZeRO-1/2 updates parameters through flat buffers and ``.data.copy_``: neither a tensor's address nor
    torch's version counter moves.  The cached bf16 / folded weight copies must follow anyway (content fingerprint,
    piggybacked on the bridge's header read; explicit check in the projector's own evaluation forward)."""
    import copy
    import ps_slm_b200.synth as S
    w, b, proj, table, br = _small_bridge(dev)
    g = torch.Generator().manual_seed(3)
    B, T = 3, 40
    raw = torch.randn(B, T + 4, 32, generator=g) * 3
    raw_lens = torch.tensor([T + 4, T - 3, T + 1])
    ids = torch.tensor([[5, 6, 299, 7], [0, 8, 299, 9], [1, 2, 299, 3]])
    mask = torch.tensor([[1, 1, 1, 1], [0, 1, 1, 1], [1, 1, 1, 1]], dtype=torch.bool)
    args = (raw.to(dev), raw_lens.to(dev), ids.to(dev), mask.to(dev))
    out0 = [t.clone() for t in br(*args) if t is not None]
    versions = [p._version for p in proj.parameters()]
    with torch.no_grad():
        for p in proj.parameters():
            p.data.copy_(p.data * 1.5 + 0.25)
        br.w_ctc.data.copy_(br.w_ctc.data.flip(0))
    assert versions == [p._version for p in proj.parameters()], "the update must be invisible to the version counter"
    out1 = [t.clone() for t in br(*args) if t is not None]
    from ps_slm_b200.bridge import TasuBridge
    fresh = TasuBridge(br.w_ctc.clone(), br.b_ctc.clone(), copy.deepcopy(proj), table.to(dev), 299, 0)
    out2 = [t for t in fresh(*args) if t is not None]
    assert not all(torch.equal(a, c) for a, c in zip(out0, out1))
    for a, c in zip(out1, out2):
        assert torch.equal(a, c), "the bridge kept using stale weight copies"
    # the projector's own (non-fused) evaluation forward
    x = torch.softmax(torch.randn(2, 7, 61, generator=g), -1).to(dev)
    y0 = proj(x).clone()
    with torch.no_grad():
        proj.ffn[0].weight.data.copy_(proj.ffn[0].weight.data * 0.5)
    y1 = proj(x)
    y2 = copy.deepcopy(proj)(x)
    assert not torch.equal(y0, y1) and torch.equal(y1, y2)


def test_audio_training_step_never_materialises_the_posterior(dev):
    """Training on an audio batch at the benchmark size (64 x 30 s): the no-grad head (TasuBridge.compress_pooled) plus the
    differentiable projector must stay far below the 3.2 GB the reference's [B, T, 25055] fp32 posterior alone takes
    (ps-slm.py:450-451), and the projected rows must equal the inference path's."""
    import ps_slm_b200.synth as S
    from ps_slm_b200.autograd import linear_silu_train_rows
    B, T = 64, 500
    w, b, proj, table, br = _bridge(dev)
    raw, raw_lens, _ = S.make_encoder_batch(B, T, w, seed=123, ragged=True)
    raw, raw_lens = raw.to(dev), raw_lens.to(dev)
    ref_rows, ref_lens, ref_max = br.compress_project(raw, raw_lens)
    ref_rows = ref_rows.float().clone()
    proj.train()
    for p in proj.parameters():
        p.requires_grad_(True)
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    pooled, mean, rstd, lens, max_len = br.compress_pooled(raw, raw_lens)
    assert torch.equal(lens, ref_lens) and max_len == ref_max and pooled.shape[0] == int(ref_lens.sum())
    y = linear_silu_train_rows(proj, pooled, mean, rstd, pooled.shape[0], torch.float32)
    y.backward(torch.randn(y.shape, device=dev, generator=torch.Generator(device=dev).manual_seed(1)))
    torch.cuda.synchronize()
    peak = torch.cuda.max_memory_allocated() - base
    posterior_bytes = B * (T + 4) * S.V_CTC * 4
    assert peak < 0.6 * posterior_bytes, (peak, posterior_bytes)
    assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in proj.parameters())
    assert ((y.detach() - ref_rows).norm() / ref_rows.norm()).item() < 1e-2


def test_bridge_fp32x3_matches_fp32_reference_to_1e5(dev):
    """north star: "posteriors, pooled features and projected embeddings within 1e-2 relative in bf16 (1e-5 in fp32)".
    ``TasuBridge.precision = "fp32x3"`` (three-term bf16 splits on the tensor cores for the kept frames' logits and both
    projector contractions, fp32 softmax / pooling / LayerNorm) against the fp32 oracle: integers bit-exact, projected
    audio rows within 1e-5 (norm-relative; every single row within 5e-5)."""
    import ps_slm_b200.synth as S
    B, T = 6, 200
    w, b, proj, table, br = _bridge(dev, torch.float32)
    br.precision = "fp32x3"
    raw, raw_lens, _ = S.make_encoder_batch(B, T, w, seed=41, ragged=True)
    ids, mask, _ = S.make_prompts(B, seed=41, left_pad=True)
    sd = {k: v.detach().cpu() for k, v in proj.state_dict().items()}
    pp = (sd["norm.weight"], sd["norm.bias"], sd["ffn.0.weight"], sd["ffn.0.bias"], sd["ffn.2.weight"], sd["ffn.2.bias"])
    with torch.no_grad():
        (e_r, m_r, _, p_r, f_r), nl_r = O.bridge_inference(raw.double(), raw_lens, w.double(), b.double(),
                                                           tuple(t.double() for t in pp), table.cpu().double(), ids, mask, None,
                                                           S.SPEECH_ID, S.PAD_ID)
    e, m, _, p, nl = br(raw.to(dev), raw_lens.to(dev), ids.to(dev), mask.to(dev))
    assert torch.equal(nl.cpu(), nl_r) and torch.equal(m.cpu(), m_r) and torch.equal(p.cpu(), p_r)
    audio = m_r & (f_r == S.PAD_ID)
    e = e.cpu().double()
    assert torch.equal(e[~audio].float(), e_r[~audio].float())
    err = ((e[audio] - e_r[audio]).norm() / e_r[audio].norm()).item()
    rows = ((e[audio] - e_r[audio]).norm(dim=-1) / e_r[audio].norm(dim=-1)).max().item()
    assert err < 1e-5 and rows < 5e-5, (err, rows)
    # the default bf16 path on the same batch, for the record (1e-2 bar)
    br.precision = "bf16"
    e16 = br(raw.to(dev), raw_lens.to(dev), ids.to(dev), mask.to(dev))[0].cpu().double()
    err16 = ((e16[audio] - e_r[audio]).norm() / e_r[audio].norm()).item()
    assert 1e-5 < err16 < 1e-2
    print("fp32x3 rel err %.2e (worst row %.2e); bf16 rel err %.2e" % (err, rows, err16))
