"""CPU tests of the host side of the token-row projector path: the batched simulator draw must be the
reference's RNG stream (ps-slm.py:380-401) and tasu_host_group_tokens a stable grouping."""
import numpy as np
import pytest
import torch

from oracle import ref_loader as R
from oracle import tasu_oracle as O

V = 25055


def _dense(tok, hot, base, lens):
    B, lmax = len(lens), max(lens) if lens else 0
    out = np.zeros((B, lmax, V), dtype=np.float32)
    r = 0
    for b, n in enumerate(lens):
        for i in range(n):
            out[b, i, :] = base[r]
            out[b, i, tok[r]] = hot[r]
            r += 1
    return out


def _transcripts(seed, B):
    g = np.random.default_rng(seed)
    ids = [g.integers(1, V, size=int(g.integers(0, 40))).tolist() for _ in range(B)]
    ids[B // 2] = []                                   # an empty transcript in the middle of the batch
    return ids


@pytest.mark.parametrize("seed", range(4))
def test_batched_draw_is_the_oracle_stream(seed):
    import ps_slm_b200.sim as sim
    ids = _transcripts(seed, 9)
    torch.manual_seed(100 + seed)
    post, lens = O.sim_posterior_noise(ids, V, 0)
    after_ref = torch.rand(1)
    torch.manual_seed(100 + seed)
    tok, hot, base, lens2 = sim.draw_noise_descriptors(ids, V, 0)
    after = torch.rand(1)
    assert lens.tolist() == lens2
    assert torch.equal(after_ref, after), "generator state after the batched draw differs"
    assert np.array_equal(_dense(tok, hot, base, lens2), post.numpy())


@pytest.mark.skipif(not R.available(), reason="reference tree not present")
def test_batched_draw_matches_reference_bitwise():
    import ps_slm_b200.sim as sim
    ids = _transcripts(7, 6)
    texts = [" ".join(map(str, i)) for i in ids]
    torch.manual_seed(5)
    a, al = R.ref_sim_noise(texts, V, 0, insert_prob=0.0)
    torch.manual_seed(5)
    tok, hot, base, lens = sim.draw_noise_descriptors([np.asarray(i, dtype=np.int32) for i in ids], V, 0)
    assert al.tolist() == lens
    assert np.array_equal(_dense(tok, hot, base, lens), a.cpu().numpy())


def test_clean_descriptors_match_oracle():
    import ps_slm_b200.sim as sim
    ids = _transcripts(3, 5)
    post, lens = O.sim_posterior_clean(ids, V)
    tok, hot, base, lens2 = sim.clean_descriptors(ids)
    assert lens.tolist() == lens2 and np.array_equal(_dense(tok, hot, base, lens2), post.numpy())


@pytest.mark.parametrize("n", [0, 1, 17, 5000])
def test_host_group_tokens(n):
    import ps_slm_b200.ops as ops
    g = np.random.default_rng(n)
    tok = g.integers(0, 300, size=n).astype(np.int32)              # heavy repetition
    hot = g.random(n, dtype=np.float32); base = g.random(n, dtype=np.float32)
    lens = [n // 2, n - n // 2]
    r = ops.group_token_rows(tok, hot, base, lens, V, "cpu")
    uq, seg, perm = r.uniq.numpy(), r.seg_off.numpy(), r.perm.numpy()
    assert r.n_rows == n and r.n_uniq == len(np.unique(tok)) and np.array_equal(uq, np.unique(tok))
    assert seg[0] == 0 and seg[-1] == n and sorted(perm.tolist()) == list(range(n))
    for u in range(r.n_uniq):
        rows = perm[seg[u]:seg[u + 1]]
        assert (tok[rows] == uq[u]).all() and (np.diff(rows) > 0).all()      # stable: ascending row order
    assert np.array_equal(r.hot.numpy(), hot) and np.array_equal(r.base.numpy(), base) and r.lens.tolist() == lens


def test_host_group_tokens_rejects_out_of_range():
    import ps_slm_b200._lib as L
    import ps_slm_b200.ops as ops
    with pytest.raises(L.TasuError):
        ops.group_token_rows(np.asarray([1, V], dtype=np.int32), np.ones(2, np.float32), np.zeros(2, np.float32), [2], V, "cpu")


def test_native_sim_call_matches_python_descriptors():
    """tasu_host_sim_token_rows (one native call) == draw_noise_descriptors + group_token_rows, same RNG stream."""
    import ps_slm_b200.ops as ops
    import ps_slm_b200.sim as sim
    for seed in range(3):
        ids = _transcripts(20 + seed, 11)
        torch.manual_seed(seed)
        a = ops.group_token_rows(*sim.draw_noise_descriptors(ids, V, 0, drop_prob=0.1, smooth_low=0.02, smooth_high=0.3),
                                 V, "cpu")
        after_a = torch.rand(1)
        torch.manual_seed(seed)
        b = ops.sim_token_rows(ops.TokenBatch(ids), V, "cpu", drop_prob=0.1, smooth_low=0.02, smooth_high=0.3)
        after_b = torch.rand(1)
        assert torch.equal(after_a, after_b) and a.n_rows == b.n_rows and a.n_uniq == b.n_uniq and a.lens_host == b.lens_host
        for k in ("uniq", "seg_off", "perm", "hot", "base", "lens"):
            assert torch.equal(getattr(a, k), getattr(b, k)), k


def test_prefetcher_preserves_draw_order():
    import ps_slm_b200.ops as ops
    import ps_slm_b200.sim as sim
    batches = [ops.TokenBatch(_transcripts(40 + i, 5)) for i in range(4)]
    torch.manual_seed(9)
    want = [ops.sim_token_rows(b, V, "cpu") for b in batches]
    torch.manual_seed(9)
    got = list(sim.TokenRowPrefetcher(batches, V, "cpu"))
    assert len(got) == 4
    for w, g in zip(want, got):
        assert torch.equal(w.perm, g.perm) and torch.equal(w.hot, g.hot) and w.lens_host == g.lens_host
