#!/usr/bin/env python
"""Default (one CTA per 128x256 tile) vs EXPERIMENTAL CTA-pair (cta_group::2, 256x256 per pair) and stream-K tcgen05 GEMM on the
deep-K shapes of the bridge: projector GEMM-1 (M = compressed rows, N = 2048, K = 25055, LN-fold + SiLU epilogue) and
GEMM-2 (N = 1536, K = 2048, bias).  CUDA events, L2 flushed between launches by a 256 MB memset.  Markdown table."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import ps_slm_b200._lib as L  # noqa: E402
import ps_slm_b200.ops as ops  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def time_ms(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    evs = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


print("| shape (M, N, K) | epilogue | one CTA per tile ms | TFLOP/s | CTA pairs ms | TFLOP/s | one-CTA stream-K ms | TFLOP/s | pair stream-K ms | TFLOP/s | pair stream-K / pairs |")
print("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for (M, N, K, epi, name) in [(8341, 2048, 25055, L.EPI_LNFOLD_SILU, "LN-fold + SiLU"), (8341, 1536, 2048, L.EPI_BIAS, "bias"),
                             (16384, 2048, 25055, L.EPI_LNFOLD_SILU, "LN-fold + SiLU"), (8192, 8192, 8192, L.EPI_NONE, "none"),
                             (11330, 25055, 512, L.EPI_SOFTMAX, "softmax (kept frames)")]:
    torch.manual_seed(0)
    A = (torch.randn(M, ops.pad_to(K), device=dev) * 0.3).bfloat16()
    B = (torch.randn(N, ops.pad_to(K), device=dev) * 0.3).bfloat16()
    C = torch.empty(M, ops.pad_to(N), dtype=torch.bfloat16, device=dev)[:, :N]      # 16-byte aligned row pitch
    bias, rstd, mean, colsum = (torch.randn(n, device=dev) for n in (N, M, M, N))
    fns = {
        "one": (0, lambda: ops.gemm_bf16_tn(A, B, M, N, K, C, epi, bias, rstd, mean, colsum)),
        "pair": (1, lambda: ops.gemm_bf16_tn(A, B, M, N, K, C, epi, bias, rstd, mean, colsum)),
        "sk_one": (0, lambda: ops.gemm_bf16_tn_streamk(A, B, M, N, K, C, epi, bias, rstd, mean, colsum)),
        "sk_pair": (1, lambda: ops.gemm_bf16_tn_streamk(A, B, M, N, K, C, epi, bias, rstd, mean, colsum)),
    }
    ts = {k: [] for k in fns}
    for rep in range(14):                                   # interleaved: clock / power drift hits every variant alike
        for k, (pair, fn) in fns.items():
            ops.set_option(L.OPT_GEMM_PAIR, pair)
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            if rep >= 2:
                ts[k].append(e0.elapsed_time(e1))
    ops.set_option(L.OPT_GEMM_PAIR, 1)
    r = {k: sorted(v)[len(v) // 2] for k, v in ts.items()}
    fl = 2.0 * M * N * K
    print(f"| {M}, {N}, {K} | {name} | {r['one']:.3f} | {fl / r['one'] / 1e9:.0f} | {r['pair']:.3f} | {fl / r['pair'] / 1e9:.0f} "
          f"| {r['sk_one']:.3f} | {fl / r['sk_one'] / 1e9:.0f} | {r['sk_pair']:.3f} | {fl / r['sk_pair'] / 1e9:.0f} | {r['sk_pair'] / r['pair']:.3f} |")
