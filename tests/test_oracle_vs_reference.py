"""CPU, build container only: the oracle restatement against the UNMODIFIED reference on random
cases (skipped where /root/reference does not exist, e.g. on the GPU box)."""
import pytest
import torch

from oracle import ref_loader as R
from oracle import tasu_oracle as O

pytestmark = pytest.mark.skipif(not R.available(), reason="reference tree not present")
SP, PAD = 99, 0


def _rand_post(g, B, T, V, logp=False):
    lab = torch.randint(0, V, (B, T), generator=g)
    lab[torch.rand(B, T, generator=g) < 0.5] = 0
    for t in range(1, T):
        rep = torch.rand(B, generator=g) < 0.4
        lab[rep, t] = lab[rep, t - 1]
    logits = torch.randn(B, T, V, generator=g)
    logits.scatter_add_(2, lab.unsqueeze(-1), (torch.rand(B, T, generator=g) * 6).unsqueeze(-1))
    p = torch.softmax(logits, -1)
    return p.log() if logp else p


@pytest.mark.parametrize("seed", range(40))
def test_psd_matches_reference(seed):
    g = torch.Generator().manual_seed(seed)
    B, T, V = (int(torch.randint(lo, hi, (1,), generator=g)) for lo, hi in ((1, 5), (1, 30), (2, 12)))
    post = _rand_post(g, B, T, V, logp=seed % 5 == 0)
    lens = torch.randint(0, T + 1, (B,), generator=g)
    feats = post if seed % 3 else torch.randn(B, T, 7, generator=g)
    a, al = R.ref_psd(feats, lens, post, 0, 0.9)
    b, bl = O.psd_loop(feats, lens, post, 0, 0.9)
    c, cl, _ = O.psd_vec(feats, lens, post, 0, 0.9)
    assert a.shape == b.shape == c.shape
    assert torch.equal(al, bl) and torch.equal(al, cl)
    assert torch.allclose(a, b, rtol=1e-6, atol=1e-7) and torch.allclose(a, c, rtol=1e-5, atol=1e-7)


def test_sim_matches_reference_bitwise():
    V = 25055
    ids = [[5, 7, 7, 100, 25054, 3, 3, 9], [1], [44, 45, 46, 47, 48, 49, 50, 51, 52, 53, 54, 55]]
    texts = [" ".join(map(str, i)) for i in ids]
    a, al = R.ref_sim_clean(texts, V)
    b, bl = O.sim_posterior_clean(ids, V)
    assert torch.equal(a, b) and torch.equal(al, bl)
    for ip in (0.0, 0.5):
        torch.manual_seed(3)
        a, al = R.ref_sim_noise(texts, V, 0, insert_prob=ip)
        torch.manual_seed(3)
        b, bl = O.sim_posterior_noise(ids, V, 0, insert_prob=ip)
        assert torch.equal(al, bl) and torch.equal(a, b)
    # the product's host-side decision code draws the same stream
    import ps_slm_b200.sim as sim
    torch.manual_seed(3)
    d1 = O.sim_noise_decisions(ids, V, 0, insert_prob=0.5)
    torch.manual_seed(3)
    d2 = sim.draw_noise_decisions(ids, 0, insert_prob=0.5)
    assert d1 == d2
    assert O.sim_row_values(0.0731, V) == sim.soft_row_values(0.0731, V)
    # the flat-descriptor fast path consumes the identical stream and yields identical rows
    import numpy as np
    for ip in (0.0, 0.5):
        torch.manual_seed(11)
        dec = sim.draw_noise_decisions(ids, 0, insert_prob=ip)
        tok, hot, base, lens, _ = sim._descriptors(dec, V, False)
        torch.manual_seed(11)
        tok2, hot2, base2, lens2 = sim.draw_noise_descriptors(ids, V, 0, insert_prob=ip)
        assert lens == lens2 and np.array_equal(tok, tok2) and np.array_equal(hot, hot2) and np.array_equal(base, base2)


@pytest.mark.parametrize("seed", range(60))
def test_merge_matches_reference(seed):
    g = torch.Generator().manual_seed(seed)
    B = int(torch.randint(1, 5, (1,), generator=g))
    S = int(torch.randint(2, 10, (1,), generator=g))
    left = bool(torch.rand(1, generator=g) < 0.5)
    ids = torch.randint(1, 50, (B, S), generator=g)
    att = torch.ones(B, S, dtype=torch.long)
    for b in range(B):
        npad = int(torch.randint(0, S - 1, (1,), generator=g)) if torch.rand(1, generator=g) < 0.6 else 0
        sp = int(torch.randint(0, S - npad, (1,), generator=g))
        if left:
            att[b, :npad] = 0; ids[b, :npad] = PAD; ids[b, npad + sp] = SP
        else:
            if npad:
                att[b, S - npad:] = 0; ids[b, S - npad:] = PAD
            ids[b, sp] = SP
    M = torch.randint(0, 6, (B,), generator=g)
    if int(M.max()) == 0:
        M[0] = 1
    emb, af = torch.randn(B, S, 4, generator=g), torch.randn(B, int(M.max()), 4, generator=g)
    lab = torch.randint(0, 50, (B, S), generator=g) if seed % 2 else None
    mask = att.bool()

    def run(fn):
        try:
            return fn(af, M, emb, ids, mask, lab, SP, PAD), None
        except Exception as e:  # noqa: BLE001
            return None, type(e).__name__
    r, rerr = run(R.ref_merge)
    o, oerr = run(O.merge)
    assert rerr == oerr
    if r is not None:
        for x, y in zip(r, o):
            assert (x is None and y is None) or (x.dtype == y.dtype and torch.equal(x, y))


@pytest.mark.parametrize("seed", range(3))
def test_cross_attention_projector_matches_reference(seed):
    g = torch.Generator().manual_seed(seed)
    V1, D, V2, B, T = 23, 48, 77, 2, 5
    ref = R.ref_projector("cross-attention", V1, D)
    post = torch.softmax(torch.randn(B, T, V1, generator=g) * 3, -1)
    table = torch.randn(V2, D, generator=g)
    with torch.no_grad():
        a = ref(post, table)
        b = O.projector_ctcca(post, table, ref.W_q.weight, ref.n_heads)
    assert a.shape == b.shape and torch.allclose(a, b, rtol=1e-6, atol=1e-7)
