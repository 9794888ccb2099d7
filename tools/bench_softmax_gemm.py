#!/usr/bin/env python
"""In-box A/B of the kept-frame softmax GEMM (M = 11330 kept frames, N = 25055, K = 512, bf16 probabilities out) under the
TEMPORARY TASU_OPT_DEBUG attribution switches (bit 0: no vector prefetch, 1: no ex2, 2: no TMA store, 3: no st.shared,
4: epilogue only drains TMEM).  Variants are interleaved round-robin so clock / power drift hits all of them alike."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import ps_slm_b200._lib as L  # noqa: E402
import ps_slm_b200.ops as ops  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
M, N, K = 11330, 25055, 512
torch.manual_seed(0)
A = (torch.randn(M, K, device=dev) * 0.3).bfloat16()
B = (torch.randn(N, K, device=dev) * 0.3).bfloat16()
C = torch.empty(M, ops.pad_to(N), dtype=torch.bfloat16, device=dev)[:, :N]
bias = torch.randn(N, device=dev)
rmax = torch.full((M,), 8.0, device=dev)
rinv = torch.full((M,), 1e-3, device=dev)
variants = [int(v) for v in (sys.argv[1:] or ["0", "1", "2", "4", "8", "12", "14", "16"])]
times = {v: [] for v in variants}
for rep in range(12):
    for v in variants:
        ops.set_option(L.OPT_DEBUG, v)
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gemm_bf16_tn(A, B, M, N, K, C, L.EPI_SOFTMAX, bias, rinv, rmax, None)
        e1.record()
        torch.cuda.synchronize()
        if rep >= 2:
            times[v].append(e0.elapsed_time(e1))
ref = None
for v in variants:
    if v in (0, 1, 64, 128, 256):                          # variants that must not change the result
        ops.set_option(L.OPT_DEBUG, v)
        C.zero_()
        ops.gemm_bf16_tn(A, B, M, N, K, C, L.EPI_SOFTMAX, bias, rinv, rmax, None)
        torch.cuda.synchronize()
        if ref is None:
            ref = C.clone()
        else:
            print("variant %d equals variant %d: %s" % (v, variants[0], bool(torch.equal(ref, C))))
ops.set_option(L.OPT_DEBUG, 0)
fl = 2.0 * M * N * K
print("| debug mask | median ms | TFLOP/s |\n|---:|---:|---:|")
for v in variants:
    t = sorted(times[v])[len(times[v]) // 2]
    print("| %d | %.4f | %.0f |" % (v, t, fl / t / 1e9))
