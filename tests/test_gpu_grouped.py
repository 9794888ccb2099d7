"""Grouped kept-frame layout (include/tasu_bridge.h step 2b', csrc/grouped.cu + the kGrouped epilogue of
csrc/gemm_sm100.cu): the mean over a run's frames (ps-slm.py:286) taken inside the kept-frame GEMM's epilogue.
Checked against the plain layout (per-frame probabilities + tasu_pool_tail), against a host model of the layout and
through the whole bridge."""
import types

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _bridge(dev):
    import ps_slm_b200.projector as P
    import ps_slm_b200.synth as S
    from ps_slm_b200.bridge import TasuBridge
    w, b = S.make_ctc_head()
    torch.manual_seed(0)
    cfg = types.SimpleNamespace(encoder_dim=S.V_CTC, llm_dim=S.H_LLM, encoder_projector_ds_rate=1)
    proj = P.EncoderProjectorLinearSiLU(cfg).to(dev).eval()
    table = S.make_embed_table(dtype=torch.bfloat16, device=dev)
    return w, b, TasuBridge(w.to(dev), b.to(dev), proj, table, S.SPEECH_ID, S.PAD_ID)


def _batch(w, B, T, seed, ragged=True, long_runs=0):
    """Planted batch; ``long_runs`` token frames are repeated over 5-9 frames (runs beyond the 4-row groups)."""
    import ps_slm_b200.synth as S
    raw, raw_lens, lab = S.make_encoder_batch(B, T, w, seed=seed, ragged=ragged)
    g = torch.Generator().manual_seed(seed + 1)
    for i in range(long_runs):
        b = i % B
        L = int(raw_lens[b]) - 4
        toks = (lab[b, :max(L - 10, 0)] != 0).nonzero().flatten()
        if toks.numel() == 0:
            continue
        t0 = int(toks[int(torch.randint(0, toks.numel(), (1,), generator=g))])
        n = 5 + i % 5
        raw[b, 4 + t0:4 + t0 + n] = raw[b, 4 + t0]
    return raw, raw_lens


def _plan(br, raw, raw_lens, dev):
    import ps_slm_b200.ops as ops
    import ps_slm_b200._lib as L
    B, T4, D = raw.shape
    T = T4 - 4
    V = br.w_ctc.shape[0]
    w_ctc, b_ctc = br._ctc_weights()
    raw_d = raw.to(dev)
    x2 = br._encoder_rows_bf16(raw_d)
    lens = torch.clamp(raw_lens.to(dev) - 4, min=0)
    st = br._head_stats(raw_d, x2, lens, w_ctc, b_ctc, B, T, D, V)
    plan = ops.collapse_plan(st, lens, 0, 0.9)
    hdr = plan.header.cpu()
    return x2, st, plan, B, T, D, V, w_ctc, b_ctc, int(hdr[L.CH_N_OUT]), int(hdr[L.CH_KEPT_FRAMES])


def _host_layout(plan, B, T):
    """The layout words and perm from the plan, restated on the host."""
    new_lens = plan.new_lens.cpu().tolist()
    seg_len = plan.seg_len.cpu().view(B, T)
    runs = [int(seg_len[b, j]) for b in range(B) for j in range(new_lens[b])]
    cls = [0 if n <= 1 else 1 if n == 2 else 2 if n <= 4 else 3 for n in runs]
    up = lambda v, a: (v + a - 1) // a * a
    ns = sum(c in (0, 3) for c in cls)
    n2, n4 = sum(c == 1 for c in cls), sum(c == 2 for c in cls)
    nxe = sum(n - 1 for n, c in zip(runs, cls) if c == 3)
    a2 = up(ns, 128)
    a4 = up(a2 + 2 * n2, 128)
    ax = up(a4 + 4 * n4, 128)
    o4 = a2 + up(n2, 64)
    ox = o4 + up(n4, 32)
    k = [0, 0, 0]
    perm = []
    for c in cls:
        if c == 1:
            perm.append(a2 + k[1]); k[1] += 1
        elif c == 2:
            perm.append(o4 + k[2]); k[2] += 1
        else:
            perm.append(k[0]); k[0] += 1
    return dict(a2=a2, a4=a4, ax=ax, a_rows=ax + nxe, o4=o4, ox=ox, n2=n2, n4=n4, ns=ns, nxe=nxe), perm, runs


@pytest.mark.parametrize("B,T,long_runs", [(6, 120, 0), (9, 300, 7), (3, 40, 2), (64, 500, 20)])
def test_grouped_pooled_rows_against_plain_layout(dev, B, T, long_runs):
    import ps_slm_b200._lib as L
    w, b, br = _bridge(dev)
    raw, raw_lens = _batch(w, B, T, seed=B * 1000 + T, long_runs=long_runs)
    x2, st, plan, B, T, D, V, w_ctc, b_ctc, n_out, n_frames = _plan(br, raw, raw_lens, dev)
    assert n_out > 0
    pooled_p, mean_p, rstd_p = br._pool_kept(x2, st, plan, B, T, D, V, n_frames, n_out, w_ctc, b_ctc)
    pooled_g, g = br._pool_kept_grouped(x2, st, plan, B, T, D, V, n_frames, n_out, w_ctc, b_ctc, n_out + 256)
    torch.cuda.synchronize()
    lay, perm, runs = _host_layout(plan, B, T)
    words = g.lay.cpu().tolist()
    got = dict(a2=words[L.GL_A2], a4=words[L.GL_A4], ax=words[L.GL_AX], a_rows=words[L.GL_A_ROWS], o4=words[L.GL_O4],
               ox=words[L.GL_OX], n2=words[L.GL_N2], n4=words[L.GL_N4], ns=words[L.GL_NS], nxe=words[L.GL_NXE])
    assert got == lay
    assert g.perm[:n_out].cpu().tolist() == perm
    if long_runs:
        assert lay["nxe"] > 0 and int(g.multi[0]) == sum(n > 4 for n in runs)
    idx = g.perm[:n_out].long()
    a, c = pooled_g[idx, :V].float(), pooled_p[:n_out, :V].float()
    single = torch.tensor([n == 1 or n > 4 for n in runs], device=dev)
    # single-frame rows and the rows of long runs go through the same arithmetic in both layouts
    assert torch.equal(a[single], c[single])
    # 2-4-frame runs: fp32 mean rounded once (grouped) vs mean of bf16-rounded frames: a few bf16 ulps of the larger one
    multi = ~single
    if bool(multi.any()):
        d = (a[multi] - c[multi]).abs()
        assert bool((d <= 2.0 ** -6 * torch.maximum(a[multi], c[multi]) + 1e-30).all())
        assert ((a[multi] - c[multi]).norm() / c[multi].norm()).item() < 4e-3
        assert (a[multi].sum(1) - 1).abs().max().item() < 2e-2          # still probability rows
    assert torch.equal(g.mean[idx], mean_p[:n_out])
    assert torch.allclose(g.rstd[idx], rstd_p[:n_out], rtol=5e-3, atol=0)
    assert torch.equal(g.rstd[idx][single], rstd_p[:n_out][single])
    # holes between the classes pool to zero rows
    holes = torch.ones(lay["ox"], dtype=torch.bool, device=dev)
    holes[idx] = False
    assert bool((pooled_g[:lay["ox"]][holes][:, :V] == 0).all())


def test_grouped_bridge_against_plain_bridge(dev):
    """Whole bridge, B = 64 x 30 s (ragged, some long runs): every integer output identical, text rows identical, audio rows
    within bf16 tolerance of the plain-layout path; two grouped calls are bit-identical."""
    import ps_slm_b200.synth as S
    B, T = 64, 500
    w, b, br = _bridge(dev)
    raw, raw_lens = _batch(w, B, T, seed=7, long_runs=12)
    ids, mask, _ = S.make_prompts(B, seed=7, left_pad=True)
    args = (raw.to(dev), raw_lens.to(dev), ids.to(dev), mask.to(dev))
    outs = {}
    for grouped in (False, True, True, True):            # calls 3, 4: speculative capacities (tail before the header)
        br.grouped_pool = grouped
        if not grouped:
            br._capacity.clear()
        e, m, _, p, nl = br(*args, want_ids=False)
        outs.setdefault(grouped, []).append((e.clone(), m.clone(), p.clone(), nl.clone()))
    (e0, m0, p0, nl0), runs = outs[False][0], outs[True]
    for e1, m1, p1, nl1 in runs:
        assert torch.equal(nl0, nl1) and torch.equal(m0, m1) and torch.equal(p0, p1)
        assert torch.equal(e1, runs[0][0]), "grouped path is not deterministic"
    e1 = runs[0][0].float()
    e0 = e0.float()
    assert ((e1 - e0).norm() / e0.norm()).item() < 3e-3
    same = (e1 == e0).all(-1)
    assert same.float().mean().item() > 0.6                # text rows, padding and single-frame audio rows are identical


def test_grouped_capacity_overflow_is_redone(dev):
    """A quiet batch followed by a dense one of the same shape: the speculative buffers are too small, nothing is written
    beyond them and the call is redone with exact sizes (same results as a fresh bridge)."""
    import ps_slm_b200.synth as S
    B, T = 32, 500
    w, b, br = _bridge(dev)
    raw, raw_lens = _batch(w, B, T, seed=11, long_runs=3)
    ids, mask, _ = S.make_prompts(B, seed=11, left_pad=True)
    args = (raw.to(dev), raw_lens.to(dev), ids.to(dev), mask.to(dev))
    ref = br(*args)
    br._capacity[(B, T)] = (64, 16)                        # high-water marks of a nearly silent batch: 2048-row buffers
    out = br(*args)
    torch.cuda.synchronize()
    for x, y in zip(ref, out):
        if x is not None:
            assert torch.equal(x, y)


def test_grouped_batch_without_kept_frames_and_with_silent_utterances(dev):
    """Edge cases of the layout: a batch of confident blanks only (no candidate at all: every region is empty) and a batch
    where some utterances keep nothing — both through the whole bridge, twice (exact-size and speculative call)."""
    import ps_slm_b200.synth as S
    B, T = 6, 90
    w, b, br = _bridge(dev)
    raw, raw_lens, lab, soft = S.make_encoder_batch(B, T, w, seed=3, ragged=True, return_soft=True)
    quiet = ((lab == 0) & ~soft).nonzero()[0]
    blank_row = raw[int(quiet[0]), 4 + int(quiet[1])].clone()
    ids, mask, _ = S.make_prompts(B, seed=3, left_pad=True)
    silent = raw.clone()
    silent[:, 4:] = blank_row
    mixed = raw.clone()
    mixed[1::2, 4:] = blank_row                              # every second utterance keeps nothing
    for batch, expect_zero in ((silent, range(B)), (mixed, range(1, B, 2))):
        outs = []
        for grouped in (False, True, True):
            br.grouped_pool = grouped
            if not grouped:
                br._capacity.clear()
            e, m, _, p, nl = br(batch.to(dev), raw_lens.to(dev), ids.to(dev), mask.to(dev))
            outs.append((e.clone(), m.clone(), p.clone(), nl.clone()))
        for bb in expect_zero:
            assert int(outs[0][3][bb]) == 0
        for e, m, p, nl in outs[1:]:
            assert torch.equal(nl, outs[0][3]) and torch.equal(m, outs[0][1]) and torch.equal(p, outs[0][2])
            assert ((e.float() - outs[0][0].float()).norm() / outs[0][0].float().norm().clamp_min(1e-9)).item() < 3e-3
        assert torch.equal(outs[1][0], outs[2][0])
