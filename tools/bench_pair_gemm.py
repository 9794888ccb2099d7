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


print("| shape (M, N, K) | epilogue | default ms | TFLOP/s | pair ms | TFLOP/s | pair / default | stream-K ms | TFLOP/s | stream-K / default |")
print("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|")
for (M, N, K, epi, name) in [(8341, 2048, 25055, L.EPI_LNFOLD_SILU, "LN-fold + SiLU"), (8341, 1536, 2048, L.EPI_BIAS, "bias"),
                             (16384, 2048, 25055, L.EPI_LNFOLD_SILU, "LN-fold + SiLU"), (8192, 8192, 8192, L.EPI_NONE, "none"),
                             (11330, 25055, 512, L.EPI_SOFTMAX, "softmax (kept frames)")]:
    torch.manual_seed(0)
    A = (torch.randn(M, ops.pad_to(K), device=dev) * 0.3).bfloat16()
    B = (torch.randn(N, ops.pad_to(K), device=dev) * 0.3).bfloat16()
    C = torch.empty(M, ops.pad_to(N), dtype=torch.bfloat16, device=dev)[:, :N]      # 16-byte aligned row pitch
    bias, rstd, mean, colsum = (torch.randn(n, device=dev) for n in (N, M, M, N))
    res = {}
    for pair in (0, 1):
        ops.set_option(L.OPT_GEMM_PAIR, 3 * pair)
        res[pair] = time_ms(lambda: ops.gemm_bf16_tn(A, B, M, N, K, C, epi, bias, rstd, mean, colsum))
    ops.set_option(L.OPT_GEMM_PAIR, 0)
    sk = time_ms(lambda: ops.gemm_bf16_tn_streamk(A, B, M, N, K, C, epi, bias, rstd, mean, colsum))
    fl = 2.0 * M * N * K
    print(f"| {M}, {N}, {K} | {name} | {res[0]:.3f} | {fl / res[0] / 1e9:.0f} | {res[1]:.3f} | {fl / res[1] / 1e9:.0f} | {res[1] / res[0]:.3f} "
          f"| {sk:.3f} | {fl / sk / 1e9:.0f} | {sk / res[0]:.3f} |")
