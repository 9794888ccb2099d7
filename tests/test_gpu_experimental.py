"""GPU tests of code paths that are written but NOT yet validated on a B200 (the round's GPU budget ran out before
they could be run).  They are skipped unless TASU_EXPERIMENTAL=1, so the default `pytest -m gpu` run only exercises the
validated path; `tools/gpu_experimental.sh` runs them under a timeout.

  * CTA-pair GEMM (tasu_set_option(TASU_OPT_GEMM_PAIR, 1)): tcgen05.mma.cta_group::2, M = 256 per pair of CTAs.
  * stream-K GEMM (tasu_gemm_bf16_tn_streamk): the ragged last wave of tiles cut along K, fix-up through a workspace.
  * epilogue prefetch (tasu_set_option(TASU_OPT_EPI_PREFETCH, mask)): bias / row vectors of the next tile fetched early.
  * 16-epilogue-warp fused CTC head (tasu_set_option(TASU_OPT_STATS_WIDE, 1)).
  * 16-independent-epilogue-warp shallow-K GEMM (tasu_set_option(TASU_OPT_GEMM_WIDE_EPI, 1)).
"""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("TASU_EXPERIMENTAL") != "1", reason="experimental paths: set TASU_EXPERIMENTAL=1")]


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


@pytest.fixture()
def pair_mode():
    import ps_slm_b200._lib as L
    import ps_slm_b200.ops as ops
    ops.set_option(L.OPT_GEMM_PAIR, 3)          # bit 0: deep-K shapes, bit 1: K <= 1024
    yield
    ops.set_option(L.OPT_GEMM_PAIR, 0)


_PROBLEM = {}


def _problem(M, N, K):
    """bf16 operands (padded pitches, garbage beyond K) and their fp64 product; the last shape is kept so that the
    epilogue variants of one shape (the fastest-varying test parameter) share the expensive reference product."""
    import ps_slm_b200.ops as ops
    key = (M, N, K)
    if key not in _PROBLEM:
        _PROBLEM.clear()
        torch.manual_seed(M + 3 * N + 7 * K)
        lda, ldb = ops.pad_to(K, 8) + 8, ops.pad_to(K, 8)
        A = torch.zeros(M, lda).bfloat16(); A[:, :K] = (torch.randn(M, K) * 0.5).bfloat16(); A[:, K:] = 7.0
        B = torch.zeros(N, ldb).bfloat16(); B[:, :K] = (torch.randn(N, K) * 0.5).bfloat16(); B[:, K:] = 7.0
        _PROBLEM[key] = (A, B, A[:, :K].float().double() @ B[:, :K].float().double().T)
    return _PROBLEM[key]


def _ref(acc, epi, bias, rstd, mean, colsum):
    if epi in (4, 5):
        z = rstd.double()[:, None] * (acc - mean.double()[:, None] * colsum.double()[None, :]) + bias.double()[None, :]
        return torch.nn.functional.silu(z) if epi == 4 else z
    if epi == 6:
        return torch.exp(acc + bias.double()[None, :] - mean.double()[:, None]) * rstd.double()[:, None]
    if epi >= 1:
        acc = acc + bias.double()[None, :]
    if epi == 2:
        acc = torch.nn.functional.silu(acc)
    if epi == 3:
        acc = torch.relu(acc)
    return acc


def _check_pad(pad, N, epi, elem_size):
    """TMA stores are clipped at N rounded up to the next 16-byte boundary of the row: pad columns inside that boundary
    receive the epilogue of a zero accumulator (zero for the linear epilogues; exp(-row_max)/sum > 0, tiny, for the
    softmax epilogue, whose bias vector is not masked there), columns beyond it are never written."""
    per16 = 16 // elem_size
    inside = (N + per16 - 1) // per16 * per16 - N
    assert bool((pad[:, inside:] == -768.0).all()), "columns beyond the 16-byte boundary must be untouched"
    head = pad[:, :inside]
    if epi == 6:
        assert bool(((head == -768.0) | ((head >= 0) & (head < 1e-3))).all())
    else:
        assert bool(((head == -768.0) | (head == 0.0)).all()), "pad columns must be untouched or zero"


# M > 128 selects the pair kernel (K > 1024: 6-stage / one epilogue group; K <= 1024: 4-stage / two groups): ragged
# M / N / K, M below / above one pair tile, a tile whose upper CTA is entirely out of range (M = 300: rows 256..299 live
# in the lower CTA of the second pair tile), the kept-frame softmax GEMM shape (K = 512, N = 25055)
PAIR_SHAPES = [(256, 256, 1088), (129, 8, 1032), (300, 260, 1100), (257, 1536, 2048), (1000, 2048, 25055), (8341, 2048, 4096),
               (256, 256, 64), (130, 260, 72), (200, 300, 1000), (900, 25055, 512)]


@pytest.mark.parametrize("epi,out_dtype", [(0, torch.float32), (1, torch.float32), (2, torch.bfloat16), (3, torch.bfloat16),
                                           (4, torch.bfloat16), (5, torch.float32), (6, torch.bfloat16)])
@pytest.mark.parametrize("M,N,K", PAIR_SHAPES)
def test_pair_gemm_matches_default_kernel(dev, pair_mode, M, N, K, epi, out_dtype):
    """The CTA-pair kernel accumulates every output element in the same order as the default kernel (one TMEM
    accumulator, K blocks in ascending order), so the two must agree BIT FOR BIT; both are checked against fp64."""
    import ps_slm_b200._lib as L
    import ps_slm_b200.ops as ops
    if (K > 20000 or N > 20000) and epi not in (1, 4, 6):
        pytest.skip("large shapes: a subset of epilogues is enough")
    A, B, acc64 = _problem(M, N, K)
    ldc = ops.pad_to(N, 8)
    torch.manual_seed(M + 3 * N + 7 * K + epi)
    bias, rstd, colsum = torch.randn(N), torch.rand(M) + 0.5, torch.randn(N)
    mean = torch.randn(M) * 0.1 if epi != 6 else torch.full((M,), 1.3 * (K ** 0.5))     # softmax: a plausible row max
    Ad, Bd = A.to(dev), B.to(dev)
    vec = [t.to(dev) for t in (bias, rstd, mean, colsum)]
    C = torch.full((M, ldc), -768.0, dtype=out_dtype, device=dev)
    assert ops.get_option(L.OPT_GEMM_PAIR) == 3
    ops.gemm_bf16_tn(Ad, Bd, M, N, K, C, epi, *vec)
    torch.cuda.synchronize()
    ops.set_option(L.OPT_GEMM_PAIR, 0)
    C0 = torch.full((M, ldc), -768.0, dtype=out_dtype, device=dev)
    ops.gemm_bf16_tn(Ad, Bd, M, N, K, C0, epi, *vec)
    torch.cuda.synchronize()
    ref = _ref(acc64, epi, bias, rstd, mean, colsum)
    got = C.cpu()
    pad = got[:, N:].float()
    _check_pad(pad, N, epi, got.element_size())
    scale = ref.abs().max().item() + 1e-6
    tol = 1e-4 if out_dtype == torch.float32 else 6e-3
    err = (got[:, :N].double() - ref).abs().max().item() / scale
    assert err < tol, f"pair GEMM max error {err} (scaled) for {(M, N, K, epi)}"
    assert torch.equal(C[:, :N], C0[:, :N]), "pair kernel and default kernel must agree bit for bit"


def test_pair_gemm_device_side_row_count(dev, pair_mode):
    """m_dev: live rows below the capacity M, including a live count that leaves whole pair tiles (and the upper CTA
    of the last live one) without work."""
    import ps_slm_b200.ops as ops
    torch.manual_seed(3)
    M, N, K = 1024, 512, 2048
    A = (torch.randn(M, K, device=dev) * 0.3).bfloat16()
    B = (torch.randn(N, K, device=dev) * 0.3).bfloat16()
    ref = A.float() @ B.float().T
    for live in (0, 1, 128, 129, 300, 1024):
        C = torch.full((M, N), -5.0, dtype=torch.float32, device=dev)
        m_dev = torch.tensor([live], dtype=torch.int32, device=dev)
        ops.gemm_bf16_tn(A, B, M, N, K, C, m_dev=m_dev)
        torch.cuda.synchronize()
        if live:
            assert (C[:live] - ref[:live]).abs().max().item() / ref.abs().max().item() < 1e-4
        tiles = (live + 255) // 256
        assert bool((C[min(M, tiles * 256):] == -5.0).all()), "rows of pair tiles without live rows must stay untouched"


def test_pair_gemm_back_to_back(dev, pair_mode):
    """Pipeline state (mbarrier phases, TMEM accumulator ring) survives many tiles per CTA pair and repeated launches."""
    import ps_slm_b200.ops as ops
    torch.manual_seed(5)
    M, N, K = 8341, 2048, 2048           # 33 x 8 = 264 pair tiles on 74 pairs: 3.6 tiles per pair
    A = (torch.randn(M, K, device=dev) * 0.3).bfloat16()
    B = (torch.randn(N, K, device=dev) * 0.3).bfloat16()
    C1 = torch.empty(M, N, dtype=torch.float32, device=dev)
    C2 = torch.empty_like(C1)
    for _ in range(3):
        ops.gemm_bf16_tn(A, B, M, N, K, C1)
    ops.gemm_bf16_tn(A, B, M, N, K, C2)
    torch.cuda.synchronize()
    assert torch.equal(C1, C2)
    ref = A.float() @ B.float().T
    assert (C1 - ref).abs().max().item() / ref.abs().max().item() < 1e-4


def test_bridge_with_pair_gemm_matches_default(dev, pair_mode):
    """Whole inference bridge with the projector GEMMs in pair mode: integers and embeddings equal the default path."""
    import types

    import ps_slm_b200._lib as L
    import ps_slm_b200.ops as ops
    import ps_slm_b200.projector as P
    import ps_slm_b200.synth as S
    from ps_slm_b200.bridge import TasuBridge
    torch.manual_seed(0)
    B, T = 8, 500
    w, b = S.make_ctc_head()
    raw, raw_lens, _ = S.make_encoder_batch(B, T, w, seed=11, ragged=True)
    ids, mask, _ = S.make_prompts(B, seed=5, left_pad=True)
    cfg = types.SimpleNamespace(encoder_dim=S.V_CTC, llm_dim=S.H_LLM, encoder_projector_ds_rate=1)
    proj = P.EncoderProjectorLinearSiLU(cfg).to(dev).eval()
    table = S.make_embed_table(dtype=torch.bfloat16, device=dev)
    br = TasuBridge(w.to(dev), b.to(dev), proj, table, S.SPEECH_ID, S.PAD_ID)
    args = (raw.to(dev), raw_lens.to(dev), ids.to(dev), mask.to(dev))
    out_pair = [t.clone() for t in br(*args)]
    torch.cuda.synchronize()
    ops.set_option(L.OPT_GEMM_PAIR, 0)
    out_def = br(*args)
    torch.cuda.synchronize()
    for a, d in zip(out_pair, out_def):
        assert torch.equal(a, d)


# ---------------------------------------------------------------------------------------------------------------
# stream-K tail (tasu_gemm_bf16_tn_streamk)
# ---------------------------------------------------------------------------------------------------------------
# (M, N, K): 528 tiles = 3 waves + 84 cut tiles (the headline GEMM-1 shape with a shorter K); fewer tiles than CTAs
# (every tile cut 4 ways); exact multiple of 148 tiles (no cut); one leftover tile; ragged M / N / K; tiny K (1 K-block)
SK_SHAPES = [(8341, 2048, 4096), (300, 512, 2048), (128 * 37, 1024, 1088), (128 * 37 + 1, 1024, 1088), (1000, 2048, 25055),
             (130, 260, 72), (257, 1536, 2048), (5, 40, 64)]


@pytest.mark.parametrize("epi,out_dtype", [(0, torch.float32), (1, torch.float32), (2, torch.bfloat16), (4, torch.bfloat16),
                                           (6, torch.bfloat16)])
@pytest.mark.parametrize("M,N,K", SK_SHAPES)
def test_streamk_gemm(dev, M, N, K, epi, out_dtype):
    import ps_slm_b200.ops as ops
    if K > 20000 and epi not in (1, 4):
        pytest.skip("large K: a subset of epilogues is enough")
    A, B, acc64 = _problem(M, N, K)
    ldc = ops.pad_to(N, 8)
    torch.manual_seed(M + 3 * N + 7 * K + epi)
    bias, rstd, colsum = torch.randn(N), torch.rand(M) + 0.5, torch.randn(N)
    mean = torch.randn(M) * 0.1 if epi != 6 else torch.full((M,), 1.3 * (K ** 0.5))
    Ad, Bd = A.to(dev), B.to(dev)
    vec = [t.to(dev) for t in (bias, rstd, mean, colsum)]
    outs = []
    for _ in range(3):                                   # the workspace flags must be handed back after every launch
        C = torch.full((M, ldc), -768.0, dtype=out_dtype, device=dev)
        ops.gemm_bf16_tn_streamk(Ad, Bd, M, N, K, C, epi, *vec)
        torch.cuda.synchronize()
        outs.append(C)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2]), "fixed summation order: deterministic"
    ref = _ref(acc64, epi, bias, rstd, mean, colsum)
    got = outs[0].cpu()
    pad = got[:, N:].float()
    _check_pad(pad, N, epi, got.element_size())
    scale = ref.abs().max().item() + 1e-6
    tol = 1e-4 if out_dtype == torch.float32 else 6e-3
    err = (got[:, :N].double() - ref).abs().max().item() / scale
    assert err < tol, f"stream-K GEMM max error {err} (scaled) for {(M, N, K, epi)}"
    assert int(ops.streamk_workspace(dev)[:4 * 148].view(torch.int32).abs().sum()) == 0, "flags are zero between launches"


def test_streamk_gemm_device_side_row_count(dev):
    import ps_slm_b200.ops as ops
    torch.manual_seed(3)
    M, N, K = 4096, 2048, 2048
    A = (torch.randn(M, K, device=dev) * 0.3).bfloat16()
    B = (torch.randn(N, K, device=dev) * 0.3).bfloat16()
    ref = A.float() @ B.float().T
    for live in (0, 1, 128, 129, 2500, 4096):            # the schedule is computed in the kernel from the live row count
        C = torch.full((M, N), -5.0, dtype=torch.float32, device=dev)
        m_dev = torch.tensor([live], dtype=torch.int32, device=dev)
        ops.gemm_bf16_tn_streamk(A, B, M, N, K, C, m_dev=m_dev)
        torch.cuda.synchronize()
        if live:
            assert (C[:live] - ref[:live]).abs().max().item() / ref.abs().max().item() < 1e-4
        tiles = (live + 127) // 128
        assert bool((C[min(M, tiles * 128):] == -5.0).all()), "rows of tiles without live rows must stay untouched"


def test_streamk_matches_default_on_uncut_tiles(dev):
    """The tiles of the full waves take exactly the default path: bit-equal to tasu_gemm_bf16_tn there."""
    import ps_slm_b200.ops as ops
    torch.manual_seed(9)
    M, N, K = 8341, 2048, 2048                           # 528 tiles: tiles 0..443 are whole, 444..527 are cut
    A = (torch.randn(M, K, device=dev) * 0.3).bfloat16()
    B = (torch.randn(N, K, device=dev) * 0.3).bfloat16()
    C0 = torch.empty(M, N, dtype=torch.float32, device=dev)
    C1 = torch.empty_like(C0)
    ops.gemm_bf16_tn(A, B, M, N, K, C0)
    ops.gemm_bf16_tn_streamk(A, B, M, N, K, C1)
    torch.cuda.synchronize()
    whole_rows = (444 // 8) * 128                        # n fastest: tile = m_tile * 8 + n_tile
    assert torch.equal(C0[:whole_rows], C1[:whole_rows])
    assert (C0 - C1).abs().max().item() / C0.abs().max().item() < 1e-5


# ---------------------------------------------------------------------------------------------------------------
# CTA-pair mode of the fused CTC head + statistics kernel (TASU_OPT_GEMM_PAIR bit 2)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,T,P,V,K,blank", [(3, 37, 4, 25055, 512, 0), (2, 130, 4, 300, 64, 7), (5, 300, 4, 4099, 512, 0),
                                             (64, 500, 4, 25055, 512, 0), (1, 200, 0, 1000, 448, 999)])
def test_pair_ctc_head_stats(dev, B, T, P, V, K, blank):
    import numpy as np

    import ps_slm_b200._lib as L
    import ps_slm_b200.ops as ops
    torch.manual_seed(V + T)
    rows = B * (T + P)
    x = (torch.randn(rows, K) * 0.7).bfloat16()
    w = (torch.randn(V, K) * 0.6).bfloat16()
    lab = torch.randint(0, V, (rows,))
    x += (4.0 * w[lab].float() / w[lab].float().norm(dim=1, keepdim=True)).bfloat16()     # a clear winner per frame
    bias = torch.randn(V) * 0.1
    xd = torch.zeros(rows, ops.pad_to(K), dtype=torch.bfloat16); xd[:, :K] = x
    wd = torch.zeros(V, ops.pad_to(K), dtype=torch.bfloat16); wd[:, :K] = w
    xd, wd, bd = xd.to(dev), wd.to(dev), bias.to(dev)
    st0 = ops.ctc_head_stats(xd, wd, bd, B, T, P, V, K, blank)
    torch.cuda.synchronize()
    ops.set_option(L.OPT_GEMM_PAIR, 4)
    try:
        st1 = ops.ctc_head_stats(xd, wd, bd, B, T, P, V, K, blank)
        st2 = ops.ctc_head_stats(xd, wd, bd, B, T, P, V, K, blank)
        torch.cuda.synchronize()
    finally:
        ops.set_option(L.OPT_GEMM_PAIR, 0)
    # logits are accumulated in the same order (one TMEM accumulator, ascending K): max / argmax / blank logit are
    # bit-equal to the default kernel; the exp-sums may be split differently over the vocabulary → fp32 rounding
    assert torch.equal(st1.argmax, st0.argmax) and torch.equal(st1.row_max, st0.row_max) and torch.equal(st1.x_blank, st0.x_blank)
    np.testing.assert_allclose(st1.row_sumexp.cpu().numpy(), st0.row_sumexp.cpu().numpy(), rtol=1e-5)
    np.testing.assert_allclose(st1.row_sumexp2.cpu().numpy(), st0.row_sumexp2.cpu().numpy(), rtol=1e-5)
    for a, b in ((st1.argmax, st2.argmax), (st1.row_sumexp, st2.row_sumexp), (st1.row_sumexp2, st2.row_sumexp2)):
        assert torch.equal(a, b), "back-to-back launches must agree bit for bit"


# ---------------------------------------------------------------------------------------------------------------
# epilogue vectors fetched one tile ahead (TASU_OPT_EPI_PREFETCH): same values, same arithmetic → bit-identical
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("epi,out_dtype", [(0, torch.float32), (1, torch.float32), (2, torch.bfloat16), (4, torch.bfloat16),
                                           (5, torch.float32), (6, torch.bfloat16)])
@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (130, 260, 72), (200, 300, 1000), (900, 25055, 512), (11330, 4099, 512), (5, 40, 136)])
def test_epilogue_prefetch_gemm_is_bit_identical(dev, M, N, K, epi, out_dtype):
    import ps_slm_b200._lib as L
    import ps_slm_b200.ops as ops
    A, B, acc64 = _problem(M, N, K)
    ldc = ops.pad_to(N, 8)
    torch.manual_seed(M + 3 * N + 7 * K + epi)
    bias, rstd, colsum = torch.randn(N), torch.rand(M) + 0.5, torch.randn(N)
    mean = torch.randn(M) * 0.1 if epi != 6 else torch.full((M,), 1.3 * (K ** 0.5))
    Ad, Bd = A.to(dev), B.to(dev)
    vec = [t.to(dev) for t in (bias, rstd, mean, colsum)]
    C0 = torch.full((M, ldc), -768.0, dtype=out_dtype, device=dev)
    ops.gemm_bf16_tn(Ad, Bd, M, N, K, C0, epi, *vec)
    torch.cuda.synchronize()
    ops.set_option(L.OPT_EPI_PREFETCH, 1)
    try:
        C1 = torch.full((M, ldc), -768.0, dtype=out_dtype, device=dev)
        ops.gemm_bf16_tn(Ad, Bd, M, N, K, C1, epi, *vec)
        torch.cuda.synchronize()
    finally:
        ops.set_option(L.OPT_EPI_PREFETCH, 0)
    assert torch.equal(C0, C1)
    ref = _ref(acc64, epi, bias, rstd, mean, colsum)
    scale = ref.abs().max().item() + 1e-6
    assert (C1.cpu()[:, :N].double() - ref).abs().max().item() / scale < (1e-4 if out_dtype == torch.float32 else 6e-3)


@pytest.mark.parametrize("B,T,P,V,K,blank", [(3, 37, 4, 25055, 512, 0), (2, 130, 4, 300, 64, 7), (64, 500, 4, 25055, 512, 0)])
def test_epilogue_prefetch_ctc_head_stats_is_bit_identical(dev, B, T, P, V, K, blank):
    import ps_slm_b200._lib as L
    import ps_slm_b200.ops as ops
    torch.manual_seed(V + T)
    rows = B * (T + P)
    xd = torch.zeros(rows, ops.pad_to(K), dtype=torch.bfloat16); xd[:, :K] = (torch.randn(rows, K) * 0.7).bfloat16()
    wd = torch.zeros(V, ops.pad_to(K), dtype=torch.bfloat16); wd[:, :K] = (torch.randn(V, K) * 0.6).bfloat16()
    xd, wd, bd = xd.to(dev), wd.to(dev), (torch.randn(V) * 0.1).to(dev)
    st0 = ops.ctc_head_stats(xd, wd, bd, B, T, P, V, K, blank)
    torch.cuda.synchronize()
    ops.set_option(L.OPT_EPI_PREFETCH, 2)
    try:
        st1 = ops.ctc_head_stats(xd, wd, bd, B, T, P, V, K, blank)
        torch.cuda.synchronize()
    finally:
        ops.set_option(L.OPT_EPI_PREFETCH, 0)
    for name in ("argmax", "x_blank", "row_max", "row_sumexp", "row_sumexp2"):
        assert torch.equal(getattr(st0, name), getattr(st1, name)), name


# ---------------------------------------------------------------------------------------------------------------
# fused CTC head + statistics with 16 epilogue warps (TASU_OPT_STATS_WIDE)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,T,P,V,K,blank", [(3, 37, 4, 25055, 512, 0), (2, 130, 4, 300, 64, 7), (1, 9, 0, 61, 32, 60),
                                             (5, 300, 4, 4099, 512, 0), (64, 500, 4, 25055, 512, 0), (2, 100, 4, 25055, 512, 25054),
                                             (2, 100, 4, 1000, 512, 200)])
def test_wide_ctc_head_stats(dev, B, T, P, V, K, blank):
    import numpy as np

    import ps_slm_b200._lib as L
    import ps_slm_b200.ops as ops
    torch.manual_seed(V + T)
    rows = B * (T + P)
    x = (torch.randn(rows, K) * 0.7).bfloat16()
    w = (torch.randn(V, K) * 0.6).bfloat16()
    lab = torch.randint(0, V, (rows,))
    x += (4.0 * w[lab].float() / w[lab].float().norm(dim=1, keepdim=True)).bfloat16()     # a clear winner per frame
    xd = torch.zeros(rows, ops.pad_to(K), dtype=torch.bfloat16); xd[:, :K] = x
    wd = torch.zeros(V, ops.pad_to(K), dtype=torch.bfloat16); wd[:, :K] = w
    xd, wd, bd = xd.to(dev), wd.to(dev), (torch.randn(V) * 0.1).to(dev)
    st0 = ops.ctc_head_stats(xd, wd, bd, B, T, P, V, K, blank)
    torch.cuda.synchronize()
    ops.set_option(L.OPT_STATS_WIDE, 1)
    try:
        st1 = ops.ctc_head_stats(xd, wd, bd, B, T, P, V, K, blank)
        st2 = ops.ctc_head_stats(xd, wd, bd, B, T, P, V, K, blank)
        torch.cuda.synchronize()
    finally:
        ops.set_option(L.OPT_STATS_WIDE, 0)
    # same logits (one accumulator, ascending K): max / argmax / blank logit bit-equal; exp-sums associated by quarters
    assert torch.equal(st1.argmax, st0.argmax) and torch.equal(st1.row_max, st0.row_max) and torch.equal(st1.x_blank, st0.x_blank)
    np.testing.assert_allclose(st1.row_sumexp.cpu().numpy(), st0.row_sumexp.cpu().numpy(), rtol=1e-5)
    np.testing.assert_allclose(st1.row_sumexp2.cpu().numpy(), st0.row_sumexp2.cpu().numpy(), rtol=1e-5)
    for a, b in ((st1.argmax, st2.argmax), (st1.row_sumexp, st2.row_sumexp), (st1.row_sumexp2, st2.row_sumexp2)):
        assert torch.equal(a, b), "back-to-back launches must agree bit for bit"


def test_bridge_with_wide_ctc_head_and_vectors_ahead_matches_default_integers(dev):
    """Whole inference bridge with the experimental epilogues: every integer output equals the default path."""
    import types

    import ps_slm_b200._lib as L
    import ps_slm_b200.ops as ops
    import ps_slm_b200.projector as P
    import ps_slm_b200.synth as S
    from ps_slm_b200.bridge import TasuBridge
    torch.manual_seed(0)
    B, T = 8, 500
    w, b = S.make_ctc_head()
    raw, raw_lens, _ = S.make_encoder_batch(B, T, w, seed=11, ragged=True)
    ids, mask, _ = S.make_prompts(B, seed=5, left_pad=True)
    cfg = types.SimpleNamespace(encoder_dim=S.V_CTC, llm_dim=S.H_LLM, encoder_projector_ds_rate=1)
    proj = P.EncoderProjectorLinearSiLU(cfg).to(dev).eval()
    table = S.make_embed_table(dtype=torch.bfloat16, device=dev)
    br = TasuBridge(w.to(dev), b.to(dev), proj, table, S.SPEECH_ID, S.PAD_ID)
    args = (raw.to(dev), raw_lens.to(dev), ids.to(dev), mask.to(dev))
    ref = [t.clone() if t is not None else None for t in br(*args)]
    torch.cuda.synchronize()
    ops.set_option(L.OPT_STATS_WIDE, 1)
    ops.set_option(L.OPT_EPI_PREFETCH, 3)
    try:
        out = br(*args)
        torch.cuda.synchronize()
    finally:
        ops.set_option(L.OPT_STATS_WIDE, 0)
        ops.set_option(L.OPT_EPI_PREFETCH, 0)
    emb, mask_o, _, pos, new_lens = out
    assert torch.equal(new_lens, ref[4]) and torch.equal(mask_o, ref[1]) and torch.equal(pos, ref[3])
    err = ((emb.float() - ref[0].float()).norm() / ref[0].float().norm()).item()
    assert err < 1e-3, err


# ---------------------------------------------------------------------------------------------------------------
# shallow-K GEMM with 16 independent epilogue warps (TASU_OPT_GEMM_WIDE_EPI): bf16 output, bit-identical
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("epi", [0, 1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 256, 64), (130, 260, 72), (200, 300, 1000), (900, 25055, 512),
                                   (11330, 4099, 512), (5, 40, 136), (1, 8, 8), (3000, 70, 512)])
def test_widegemm_epilogue_is_bit_identical(dev, M, N, K, epi):
    import ps_slm_b200._lib as L
    import ps_slm_b200.ops as ops
    A, B, acc64 = _problem(M, N, K)
    ldc = ops.pad_to(N, 8) + 8
    torch.manual_seed(M + 3 * N + 7 * K + epi)
    bias, rstd, colsum = torch.randn(N), torch.rand(M) + 0.5, torch.randn(N)
    mean = torch.randn(M) * 0.1 if epi != 6 else torch.full((M,), 1.3 * (K ** 0.5))
    Ad, Bd = A.to(dev), B.to(dev)
    vec = [t.to(dev) for t in (bias, rstd, mean, colsum)]
    C0 = torch.full((M, ldc), -768.0, dtype=torch.bfloat16, device=dev)
    ops.gemm_bf16_tn(Ad, Bd, M, N, K, C0, epi, *vec)
    torch.cuda.synchronize()
    ops.set_option(L.OPT_GEMM_WIDE_EPI, 1)
    try:
        outs = []
        for _ in range(2):
            C1 = torch.full((M, ldc), -768.0, dtype=torch.bfloat16, device=dev)
            ops.gemm_bf16_tn(Ad, Bd, M, N, K, C1, epi, *vec)
            torch.cuda.synchronize()
            outs.append(C1)
    finally:
        ops.set_option(L.OPT_GEMM_WIDE_EPI, 0)
    assert torch.equal(outs[0], outs[1])
    assert torch.equal(C0[:, :N], outs[0][:, :N]), "same arithmetic per element: bit-identical to the default kernel"
    pad = outs[0][:, N:].float().cpu()
    _check_pad(pad, N, epi, 2)
    ref = _ref(acc64, epi, bias, rstd, mean, colsum)
    scale = ref.abs().max().item() + 1e-6
    assert (outs[0].cpu()[:, :N].double() - ref).abs().max().item() / scale < 6e-3


def test_widegemm_device_side_row_count(dev):
    import ps_slm_b200._lib as L
    import ps_slm_b200.ops as ops
    torch.manual_seed(3)
    M, N, K = 2048, 1000, 512
    A = (torch.randn(M, K, device=dev) * 0.3).bfloat16()
    B = (torch.randn(N, K, device=dev) * 0.3).bfloat16()
    ref = A.float() @ B.float().T
    ops.set_option(L.OPT_GEMM_WIDE_EPI, 1)
    try:
        for live in (0, 1, 128, 129, 700, 2048):
            C = torch.full((M, ops.pad_to(N, 8)), -5.0, dtype=torch.bfloat16, device=dev)
            m_dev = torch.tensor([live], dtype=torch.int32, device=dev)
            ops.gemm_bf16_tn(A, B, M, N, K, C, m_dev=m_dev)
            torch.cuda.synchronize()
            if live:
                assert (C[:live, :N].float() - ref[:live]).abs().max().item() / ref.abs().max().item() < 6e-3
            tiles = (live + 127) // 128
            assert bool((C[min(M, tiles * 128):] == -5.0).all()), "rows of tiles without live rows must stay untouched"
    finally:
        ops.set_option(L.OPT_GEMM_WIDE_EPI, 0)


def test_bridge_bf16_encoder_output_equals_fp32_input(dev):
    """bench.py --host-bf16: handing the encoder output over as bf16 skips the cast kernel; the kernels see the same
    bf16 values, so every output is bit-identical to the fp32-input call."""
    import types

    import ps_slm_b200.projector as P
    import ps_slm_b200.synth as S
    from ps_slm_b200.bridge import TasuBridge
    torch.manual_seed(0)
    B, T = 6, 300
    w, b = S.make_ctc_head()
    raw, raw_lens, _ = S.make_encoder_batch(B, T, w, seed=21, ragged=True)
    ids, mask, _ = S.make_prompts(B, seed=7, left_pad=True)
    cfg = types.SimpleNamespace(encoder_dim=S.V_CTC, llm_dim=S.H_LLM, encoder_projector_ds_rate=1)
    proj = P.EncoderProjectorLinearSiLU(cfg).to(dev).eval()
    table = S.make_embed_table(dtype=torch.bfloat16, device=dev)
    br = TasuBridge(w.to(dev), b.to(dev), proj, table, S.SPEECH_ID, S.PAD_ID)
    tail = (raw_lens.to(dev), ids.to(dev), mask.to(dev))
    out32 = [t.clone() if t is not None else None for t in br(raw.to(dev), *tail)]
    out16 = br(raw.bfloat16().to(dev), *tail)
    torch.cuda.synchronize()
    for a, c in zip(out32, out16):
        assert (a is None and c is None) or torch.equal(a, c)
