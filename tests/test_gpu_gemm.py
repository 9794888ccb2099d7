"""GPU parity tests for the tcgen05/TMEM/TMA GEMM, the projector plugins and the fused bridge."""
import types

import numpy as np
import pytest
import torch

from oracle import tasu_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def _t(a):
    return torch.from_numpy(np.asarray(a))


def _ref_gemm(A, B, epi, bias, rstd, mean, colsum):
    acc = A.float().double() @ B.float().double().T
    if epi == 4:
        z = rstd.double()[:, None] * (acc - mean.double()[:, None] * colsum.double()[None, :]) + bias.double()[None, :]
        return torch.nn.functional.silu(z)
    if epi >= 1:
        acc = acc + bias.double()[None, :]
    if epi == 2:
        acc = torch.nn.functional.silu(acc)
    if epi == 3:
        acc = torch.relu(acc)
    return acc


SHAPES = [
    (128, 256, 64), (1, 8, 8), (130, 260, 72), (200, 300, 1000), (257, 1536, 2048), (1000, 2048, 25055),
    (900, 25055, 512), (5, 40, 136),
]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("epi,out_dtype", [(0, torch.float32), (1, torch.float32), (2, torch.bfloat16),
                                           (3, torch.bfloat16), (4, torch.bfloat16)])
def test_gemm_tcgen05(dev, M, N, K, epi, out_dtype):
    import ps_slm_b200.ops as ops
    torch.manual_seed(M * 7 + N * 3 + K + epi)
    lda, ldb = ops.pad_to(K, 8) + 8, ops.pad_to(K, 8)
    ldc = ops.pad_to(N, 8)
    A = torch.zeros(M, lda).bfloat16(); A[:, :K] = (torch.randn(M, K) * 0.5).bfloat16()
    A[:, K:] = 7.0                                     # garbage beyond K must never be read
    B = torch.zeros(N, ldb).bfloat16(); B[:, :K] = (torch.randn(N, K) * 0.5).bfloat16()
    B[:, K:] = 7.0
    bias, rstd, mean, colsum = torch.randn(N), torch.rand(M) + 0.5, torch.randn(M) * 0.1, torch.randn(N)
    C = torch.full((M, ldc), -777.0, dtype=out_dtype, device=dev)
    Ad, Bd = A.to(dev), B.to(dev)
    ops.gemm_bf16_tn(Ad, Bd, M, N, K, C, epi, bias.to(dev), rstd.to(dev), mean.to(dev), colsum.to(dev))
    torch.cuda.synchronize()
    ref = _ref_gemm(A[:, :K], B[:, :K], epi, bias, rstd, mean, colsum)
    got = C.cpu()
    # TMA stores are clipped at N up to the next 16-byte boundary of the row: pad columns inside the
    # pitch are either untouched or zero (never garbage), and nothing is written past the pitch
    pad = got[:, N:].float()
    assert bool(((pad == -777.0) | (pad == 0.0)).all()), "pad columns must be untouched or zero"
    scale = ref.abs().max().item() + 1e-6
    tol = 1e-4 if out_dtype == torch.float32 else 6e-3
    err = (got[:, :N].double() - ref).abs().max().item() / scale
    assert err < tol, f"tcgen05 GEMM max error {err} (scaled) for {(M, N, K, epi)}"
    # cross-check against the CUDA-core kernel of the same contract
    C2 = torch.empty_like(C)
    ops.gemm_bf16_tn(Ad, Bd, M, N, K, C2, epi, bias.to(dev), rstd.to(dev), mean.to(dev), colsum.to(dev), simt=True)
    err2 = (C2.cpu()[:, :N].double() - ref).abs().max().item() / scale
    assert err2 < tol


def test_gemm_back_to_back_is_deterministic(dev):
    import ps_slm_b200.ops as ops
    torch.manual_seed(5)
    M, N, K = 3000, 2048, 4096
    A = (torch.randn(M, K, device=dev) * 0.3).bfloat16()
    B = (torch.randn(N, K, device=dev) * 0.3).bfloat16()
    C1 = torch.empty(M, N, dtype=torch.float32, device=dev)
    C2 = torch.empty_like(C1)
    for _ in range(3):
        ops.gemm_bf16_tn(A, B, M, N, K, C1)
    ops.gemm_bf16_tn(A, B, M, N, K, C2)
    assert torch.equal(C1, C2)
    ref = A.float() @ B.float().T
    assert (C1 - ref).abs().max().item() / ref.abs().max().item() < 1e-4


@pytest.mark.parametrize("B,T,P,V,K,blank", [(3, 37, 4, 25055, 512, 0), (2, 130, 4, 300, 64, 7), (1, 9, 0, 61, 32, 60),
                                             (5, 300, 4, 4099, 512, 0)])
def test_ctc_head_stats_fused(dev, B, T, P, V, K, blank):
    """Fused CTC head + softmax statistics (logits never leave TMEM) vs fp32 math on the same bf16 operands."""
    import ps_slm_b200.ops as ops
    torch.manual_seed(V + T)
    x = (torch.randn(B * (T + P), K) * 0.7).bfloat16()
    w = (torch.randn(V, K) * 0.6).bfloat16()
    lab = torch.randint(0, V, (B * (T + P),))
    x += (4.0 * w[lab].float() / w[lab].float().norm(dim=1, keepdim=True)).bfloat16()     # a clear winner per frame
    bias = torch.randn(V) * 0.1
    ldx, ldw = ops.pad_to(K), ops.pad_to(K)
    xd = torch.zeros(B * (T + P), ldx, dtype=torch.bfloat16); xd[:, :K] = x
    wd = torch.zeros(V, ldw, dtype=torch.bfloat16); wd[:, :K] = w
    st = ops.ctc_head_stats(xd.to(dev), wd.to(dev), bias.to(dev), B, T, P, V, K, blank)
    torch.cuda.synchronize()
    logits = (x.double() @ w.double().T + bias.double()).view(B, T + P, V)[:, P:, :].reshape(B * T, V)
    ref_max, ref_arg = logits.max(-1)
    top2 = logits.topk(2, -1).values
    clear = (top2[:, 0] - top2[:, 1]) > 1e-3                       # ties within accumulation noise are not comparable
    assert torch.equal(st.argmax.cpu().long()[clear], ref_arg[clear])
    np.testing.assert_allclose(st.row_max.cpu().double().numpy(), ref_max.numpy(), rtol=1e-4, atol=1e-4)
    ref_sum = torch.exp(logits - ref_max[:, None]).sum(-1)
    np.testing.assert_allclose(st.row_sumexp.cpu().double().numpy(), ref_sum.numpy(), rtol=2e-3)
    np.testing.assert_allclose(st.x_blank.cpu().double().numpy(), logits[:, blank].numpy(), rtol=1e-4, atol=1e-4)
    ref_sum2 = torch.exp(2 * (logits - ref_max[:, None])).sum(-1)
    np.testing.assert_allclose(st.row_sumexp2.cpu().double().numpy(), ref_sum2.numpy(), rtol=2e-3)


def _cfg(D, H, k=1):
    return types.SimpleNamespace(encoder_dim=D, llm_dim=H, encoder_projector_ds_rate=k)


def test_projector_golden(dev, golden):
    import ps_slm_b200.projector as P
    g = golden["projector"]
    for kind, cls, k, key in (("linear-silu", P.EncoderProjectorLinearSiLU, 1, "linear_silu"),
                              ("linear", P.EncoderProjectorConcat, 2, "linear"),
                              ("simple_linear", P.EncoderProjectorLinear, 3, "simple_linear")):
        x = _t(g[f"{key}_x"])
        D = x.shape[-1]
        m = cls(_cfg(D, 32, k))
        sd = {n[len(key) + 3:]: _t(g[n]) for n in g.files if n.startswith(key + "_p_")}
        missing, unexpected = m.load_state_dict(sd, strict=True)          # same parameter names as the reference
        assert not missing and not unexpected and m.k == k
        m = m.to(dev).eval()
        with torch.no_grad():
            y = m(x.to(dev)).cpu()
        ref = _t(g[f"{key}_ref_y"])
        assert y.shape == ref.shape and y.dtype == torch.float32
        err = (y - ref).norm() / ref.norm()
        assert err < 1e-2, f"{kind}: relative error {err}"           # bf16 tolerance of the north star


def test_projector_linear_silu_full_width(dev):
    """V=25055 → 2048 → 1536 on peaky posterior rows (plus a zero pad row) vs the fp32 oracle."""
    import ps_slm_b200.projector as P
    torch.manual_seed(0)
    m = P.EncoderProjectorLinearSiLU(_cfg(25055, 1536))
    with torch.no_grad():
        m.norm.weight.uniform_(0.8, 1.2); m.norm.bias.uniform_(-0.05, 0.05); m.ffn[2].bias.uniform_(-0.1, 0.1)
    x = torch.softmax(torch.randn(2, 40, 25055) * 6, -1)
    x[1, 30:] = 0
    sd = m.state_dict()
    ref = O.projector_linear_silu(x, sd["norm.weight"], sd["norm.bias"], sd["ffn.0.weight"], sd["ffn.0.bias"],
                                  sd["ffn.2.weight"], sd["ffn.2.bias"])
    m = m.to(dev).eval()
    with torch.no_grad():
        y = m(x.to(dev)).cpu()
    err = (y - ref).norm() / ref.norm()
    assert err < 1e-2, f"relative error {err}"
    rowerr = ((y - ref).norm(dim=-1) / ref.norm(dim=-1)).max()
    assert rowerr < 2e-2, f"worst row relative error {rowerr}"


def test_projector_fp32_accurate_mode(dev):
    """precision="fp32x3": three-term bf16 split on the tensor cores reproduces the reference's fp32 projector
    to ~1e-5 (north-star fp32 tolerance) at full width."""
    import ps_slm_b200.projector as P
    torch.manual_seed(0)
    m = P.EncoderProjectorLinearSiLU(_cfg(25055, 1536))
    with torch.no_grad():
        m.norm.weight.uniform_(0.8, 1.2); m.norm.bias.uniform_(-0.05, 0.05); m.ffn[2].bias.uniform_(-0.1, 0.1)
    x = torch.softmax(torch.randn(2, 24, 25055) * 6, -1)
    x[1, 20:] = 0
    sd = {k: v.double() for k, v in m.state_dict().items()}
    ref = O.projector_linear_silu(x.double(), sd["norm.weight"], sd["norm.bias"], sd["ffn.0.weight"], sd["ffn.0.bias"],
                                  sd["ffn.2.weight"], sd["ffn.2.bias"])           # fp64 ground truth
    ref32 = O.projector_linear_silu(x, *[m.state_dict()[k] for k in ("norm.weight", "norm.bias", "ffn.0.weight",
                                                                     "ffn.0.bias", "ffn.2.weight", "ffn.2.bias")])
    m = m.to(dev).eval()
    m.precision = "fp32x3"
    with torch.no_grad():
        y = m(x.to(dev)).cpu()
    err = ((y.double() - ref).norm() / ref.norm()).item()
    err_ref32 = ((ref32.double() - ref).norm() / ref.norm()).item()     # what plain fp32 on the CPU achieves
    assert err < 1e-5, f"fp32x3 relative error {err} (torch fp32 CPU: {err_ref32})"
    assert ((y - ref32).norm() / ref32.norm()).item() < 1e-5


@pytest.mark.parametrize("materialize", [False, True])
@pytest.mark.parametrize("ragged,labels", [(False, False), (True, True)])
def test_bridge_inference_vs_oracle(dev, ragged, labels, materialize):
    """Whole fused path (ctc_lo → softmax/argmax → collapse → pool → projector → splice) against the
    fp32 oracle on planted-label input: every integer bit-exact, embeddings within 1e-2."""
    import ps_slm_b200.projector as P
    import ps_slm_b200.synth as S
    from ps_slm_b200.bridge import TasuBridge
    torch.manual_seed(1)
    B, T = 5, 120
    w, b = S.make_ctc_head()
    raw, raw_lens, _ = S.make_encoder_batch(B, T, w, seed=7, ragged=ragged)
    if labels:
        ids, mask, lab = S.make_prompts(B, seed=3, left_pad=False, target_lens=[5, 9, 1, 30, 12])
    else:
        ids, mask, lab = S.make_prompts(B, seed=3, left_pad=True)
    proj = P.EncoderProjectorLinearSiLU(_cfg(S.V_CTC, S.H_LLM))
    with torch.no_grad():
        proj.norm.weight.uniform_(0.8, 1.2); proj.norm.bias.uniform_(-0.05, 0.05)
    table = S.make_embed_table(dtype=torch.float32)
    sd = proj.state_dict()
    pp = (sd["norm.weight"], sd["norm.bias"], sd["ffn.0.weight"], sd["ffn.0.bias"], sd["ffn.2.weight"], sd["ffn.2.bias"])
    (e_r, m_r, l_r, p_r, f_r), nl_r = O.bridge_inference(raw, raw_lens, w, b, pp, table, ids, mask, lab,
                                                        S.SPEECH_ID, S.PAD_ID)
    proj = proj.to(dev).eval()
    br = TasuBridge(w.to(dev), b.to(dev), proj, table.to(dev), S.SPEECH_ID, S.PAD_ID)
    br.materialize_logits = materialize           # False: fused stats + recompute of the kept frames (default)
    e, m, l, p, nl = br(raw.to(dev), raw_lens.to(dev), ids.to(dev), mask.to(dev), None if lab is None else lab.to(dev))
    assert torch.equal(nl.cpu(), nl_r)
    assert e.shape == e_r.shape
    assert torch.equal(m.cpu(), m_r) and torch.equal(p.cpu(), p_r)
    if lab is None:
        assert l is None
    else:
        assert torch.equal(l.cpu(), l_r)
    e = e.cpu().float()
    err = (e - e_r).norm() / e_r.norm()
    assert err < 1e-2, f"inputs_embeds relative error {err}"
    # audio rows alone (text rows are exact copies and would dilute the norm)
    audio = m_r & (f_r == S.PAD_ID)
    aerr = (e[audio] - e_r[audio]).norm() / e_r[audio].norm()
    assert aerr < 1e-2, f"audio embedding relative error {aerr}"
    assert torch.equal(e[~audio], e_r[~audio])


@pytest.mark.parametrize("a_mn,b_mn", [(True, True), (False, True), (True, False)])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (200, 300, 520), (1536, 2048, 1737), (777, 2048, 1536), (64, 72, 3000)])
def test_gemm_mn_major_operands(dev, a_mn, b_mn, M, N, K):
    """tasu_gemm_bf16_f32: either operand stored [K, MN] (MN contiguous) — the projector's backward contractions
    (dW2 = dy^T·h, dh = dy·W2) without materialised transposes — against the fp32 product of the same bf16 values."""
    import ps_slm_b200.ops as ops
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, generator=g).to(torch.bfloat16)          # logical A [M, K], B [N, K]
    Bm = torch.randn(N, K, generator=g).to(torch.bfloat16)
    ref = A.float() @ Bm.float().t()
    p8 = lambda n: (n + 7) // 8 * 8                                # noqa: E731
    if a_mn:                                                       # stored [K, M], padded pitch
        a_store = torch.zeros(K, p8(M), dtype=torch.bfloat16)
        a_store[:, :M] = A.t()
        a_dev = a_store.to(dev)[:, :M]
    else:
        a_store = torch.zeros(M, p8(K), dtype=torch.bfloat16)
        a_store[:, :K] = A
        a_dev = a_store.to(dev)[:, :K]
    if b_mn:
        b_store = torch.zeros(K, p8(N), dtype=torch.bfloat16)
        b_store[:, :N] = Bm.t()
        b_dev = b_store.to(dev)[:, :N]
    else:
        b_store = torch.zeros(N, p8(K), dtype=torch.bfloat16)
        b_store[:, :K] = Bm
        b_dev = b_store.to(dev)[:, :K]
    out = ops.gemm_bf16_f32(a_dev, a_mn, b_dev, b_mn, M, N, K)
    err = ((out.cpu() - ref).norm() / ref.norm()).item()
    assert err < 1e-3, f"relative error {err}"                    # exact bf16 products, (truncating) fp32 accumulation


@pytest.mark.parametrize("V1,D,V2,B,T", [(300, 64, 1000, 2, 37), (61, 48, 200, 3, 20), (25055, 1536, 4099, 1, 150)])
def test_cross_attention_projector(dev, V1, D, V2, B, T):
    """EncoderProjectorCTCCA (projector.py:104-126) composed from the stats / softmax / MN-major GEMM kernels vs the
    fp32 oracle; d = D/8 = 6 exercises the head-padding branch, d = 192 is the Qwen2.5-1.5B shape."""
    import ps_slm_b200.projector as P
    torch.manual_seed(V1 + D)
    m = P.EncoderProjectorCTCCA(types.SimpleNamespace(encoder_dim=V1, llm_dim=D, encoder_projector_ds_rate=1))
    post = torch.softmax(torch.randn(B, T, V1) * 4, -1)
    post[0, T - 3:] = 0                                              # zero-padded rows as psd() emits them
    table = torch.randn(V2, D) * 0.5
    with torch.no_grad():
        ref = O.projector_ctcca(post, table, m.W_q.weight, m.n_heads)
        md = m.to(dev).eval()
        out = md(post.to(dev), table.to(dev))
        assert out.shape == ref.shape and out.dtype == torch.float32
        err = ((out.cpu() - ref).norm() / ref.norm()).item()
        assert err < 1e-2, f"relative error {err}"
        out_bf = md(post.to(dev), table.to(dev).bfloat16())         # bf16 LLM embedding table
        assert ((out_bf.cpu() - ref).norm() / ref.norm()).item() < 2e-2


@pytest.mark.parametrize("V1,D,V2,B,T", [(300, 64, 1000, 2, 37), (61, 48, 200, 3, 20)])
def test_cross_attention_projector_backward(dev, V1, D, V2, B, T):
    """W_q gradient of the cross-attention projector (probabilities recomputed per head, dS = P∘(dP − dZ·Z)) vs torch
    autograd through the fp32 oracle."""
    import ps_slm_b200.projector as P
    torch.manual_seed(V1 + D + 1)
    m = P.EncoderProjectorCTCCA(types.SimpleNamespace(encoder_dim=V1, llm_dim=D, encoder_projector_ds_rate=1))
    post = torch.softmax(torch.randn(B, T, V1) * 4, -1)
    table = torch.randn(V2, D) * 0.5
    gz = torch.randn(B, T, D)
    wq = m.W_q.weight.detach().clone().requires_grad_(True)
    ref = O.projector_ctcca(post, table, wq, m.n_heads)
    (ref * gz).sum().backward()
    md = m.to(dev).train()
    out = md(post.to(dev), table.to(dev))
    assert out.requires_grad
    assert ((out.detach().cpu() - ref.detach()).norm() / ref.detach().norm()).item() < 1e-2
    (out * gz.to(dev)).sum().backward()
    g = md.W_q.weight.grad
    assert g is not None and g.shape == wq.grad.shape
    err = ((g.cpu() - wq.grad).norm() / wq.grad.norm()).item()
    assert err < 3e-2, f"relative gradient error {err}"


@pytest.mark.parametrize("V1,D,V2,B,T", [(200, 512, 1000, 2, 100), (64, 1024, 515, 3, 90), (300, 1536, 4099, 2, 200),
                                         (128, 2048, 300, 1, 129), (96, 1536, 151936, 1, 130)])
def test_cross_attention_fused_kernel(dev, V1, D, V2, B, T):
    """tasu_attn_softmax_pv (all heads in one launch, probabilities kept in shared memory) against the composed path it
    replaces (softmax GEMM + MN-major GEMM per head) and against the fp32 oracle (projector.py:104-126): head widths
    64 / 128 / 192 / 256, ragged last query tile and last key tile, several items per CTA, the full 151936-row table."""
    import ps_slm_b200.projector as P
    torch.manual_seed(V1 + D + V2)
    m = P.EncoderProjectorCTCCA(types.SimpleNamespace(encoder_dim=V1, llm_dim=D, encoder_projector_ds_rate=1))
    assert (D // m.n_heads) in (64, 128, 192, 256)
    post = torch.softmax(torch.randn(B, T, V1) * 4, -1)
    post[0, T - 3:] = 0
    table = (torch.randn(V2, D) * 0.5).bfloat16()
    md = m.to(dev).eval()
    saved = P.FUSED_ATTENTION
    try:
        with torch.no_grad():
            P.FUSED_ATTENTION = False
            composed = md(post.to(dev), table.to(dev))
            P.FUSED_ATTENTION = True
            fused = md(post.to(dev), table.to(dev))
            fused2 = md(post.to(dev), table.to(dev))
        torch.cuda.synchronize()
    finally:
        P.FUSED_ATTENTION = saved
    assert torch.equal(fused, fused2), "the fused kernel must be deterministic"
    scale = composed.abs().max().item() + 1e-9
    # both round the probabilities to bf16 — the composed path after the normalisation, the fused kernel before it
    assert (fused - composed).abs().max().item() / scale < 8e-3
    assert ((fused - composed).norm() / composed.norm()).item() < 4e-3
    if V2 <= 8192:
        with torch.no_grad():
            ref = O.projector_ctcca(post, table.float(), m.W_q.weight.detach().cpu(), m.n_heads)
        assert ((fused.cpu() - ref).norm() / ref.norm()).item() < 2e-2


@pytest.mark.parametrize("N,V2,h,d", [(300, 1000, 8, 192), (129, 515, 4, 64), (2000, 4099, 2, 128), (128, 256, 3, 256), (50, 77, 8, 192)])
def test_attn_softmax_pv_both_modes(dev, N, V2, h, d):
    """tasu_attn_softmax_pv against a torch fp32 attention over the same bf16 operands: with the row statistics of a
    separate pass (one sweep, normalised probabilities) and self-contained (row maxima in a first sweep, O divided by the
    fp32 sum of the probabilities at the end)."""
    import ps_slm_b200.ops as ops
    torch.manual_seed(N + V2)
    Q = (torch.randn(N, h * d, device=dev) * 0.4).bfloat16()
    table = (torch.randn(V2, h * d, device=dev) * 0.5).bfloat16()
    q = Q.float().view(N, h, d).transpose(0, 1)                          # [h, N, d]
    k = table.float().view(V2, h, d).transpose(0, 1)                     # [h, V2, d]
    s = q @ k.transpose(1, 2)
    ref = (torch.softmax(s, -1) @ k).transpose(0, 1).reshape(N, h * d)
    row_max = s.max(-1).values.contiguous()
    row_inv = (1.0 / torch.exp(s - row_max[..., None]).sum(-1)).contiguous()
    Z1 = torch.full((N, h * d), 7.0, device=dev)
    Z2 = torch.full((N, h * d), 7.0, device=dev)
    ops.attn_softmax_pv(Q, table, N, V2, h, d, Z1, row_max, row_inv)
    ops.attn_softmax_pv(Q, table, N, V2, h, d, Z2)
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    assert (Z1 - ref).abs().max().item() / scale < 6e-3                  # bf16 probabilities
    assert (Z2 - ref).abs().max().item() / scale < 6e-3
    assert (Z1 - Z2).abs().max().item() / scale < 3e-3


@pytest.mark.parametrize("N,V2,h,d", [(300, 1000, 8, 192), (2000, 4099, 2, 128), (129, 515, 4, 64), (1100, 151936, 8, 192)])
def test_attn_key_split_against_one_item_per_tile(dev, N, V2, h, d):
    """Key split of the self-contained mode (tasu_attn_softmax_pv_ws + merge kernel: every (row tile, head) item cut into
    key ranges so that the last wave is full) against the unsplit kernel: same result up to the bf16 rounding of the
    probabilities relative to different maxima, deterministic, and the plan is what the host function says."""
    import ctypes
    import ps_slm_b200.ops as ops
    import ps_slm_b200._lib as L
    torch.manual_seed(N + V2 + 1)
    Q = (torch.randn(N, h * d, device=dev) * 0.4).bfloat16()
    table = (torch.randn(V2, h * d, device=dev) * 0.5).bfloat16()
    n_splits = ctypes.c_int(0)
    ws_bytes = int(L.lib().tasu_attn_split_plan(N, V2, h, d, ctypes.byref(n_splits)))
    items = (N + 127) // 128 * h
    assert 1 <= n_splits.value <= 8 and n_splits.value <= (V2 + 127) // 128
    assert ws_bytes == (items * n_splits.value * 128 * (d + 2) * 4 if n_splits.value > 1 else 0)
    outs = []
    saved = ops.ATTN_KEY_SPLIT
    try:
        for split in (False, True, True):
            ops.ATTN_KEY_SPLIT = split
            Z = torch.full((N, h * d), 7.0, device=dev)
            ops.attn_softmax_pv(Q, table, N, V2, h, d, Z)
            outs.append(Z)
        torch.cuda.synchronize()
    finally:
        ops.ATTN_KEY_SPLIT = saved
    assert torch.equal(outs[1], outs[2]), "the key split must be deterministic"
    scale = outs[0].abs().max().item()
    assert (outs[1] - outs[0]).abs().max().item() / scale < 4e-3
    assert ((outs[1] - outs[0]).norm() / outs[0].norm()).item() < 2e-3
    if V2 <= 8192:
        q = Q.float().view(N, h, d).transpose(0, 1)
        k = table.float().view(V2, h, d).transpose(0, 1)
        ref = (torch.softmax(q @ k.transpose(1, 2), -1) @ k).transpose(0, 1).reshape(N, h * d)
        assert (outs[1] - ref).abs().max().item() / ref.abs().max().item() < 6e-3
