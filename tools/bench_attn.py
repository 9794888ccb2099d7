#!/usr/bin/env python
"""Cross-attention projector (projector.py:104-126) on a config-2-sized compressed batch ([64, 154, 25055] posterior rows
against the 151 936-row Qwen2.5-1.5B-shaped table): the fused kernel (tasu_attn_softmax_pv) against the composed path
(softmax GEMM + MN-major GEMM per head), variants interleaved in one process, CUDA events."""
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ps_slm_b200.ops as ops
import ps_slm_b200.projector as P
import ps_slm_b200.synth as S


def main():
    dev = torch.device("cuda", 0)
    B, T, V = 64, 154, S.V_CTC
    torch.manual_seed(0)
    post = torch.softmax(torch.randn(B, T, V, device=dev) * 3, -1)
    table = S.make_embed_table(dtype=torch.bfloat16, device=dev)
    cfg = types.SimpleNamespace(encoder_dim=V, llm_dim=S.H_LLM, encoder_projector_ds_rate=1)
    ca = P.EncoderProjectorCTCCA(cfg).to(dev).eval()
    N, h, V2, D = B * T, ca.n_heads, table.shape[0], S.H_LLM
    d = D // h
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def run(fused):
        P.FUSED_ATTENTION = fused
        with torch.no_grad():
            return ca(post, table)

    res = {True: [], False: []}
    outs = {}
    for rep in range(4):
        for fused in (False, True):
            torch.cuda.synchronize()
            e0.record(); outs[fused] = run(fused); e1.record()
            torch.cuda.synchronize()
            if rep:
                res[fused].append(e0.elapsed_time(e1))
    err = (outs[True] - outs[False]).abs().max().item() / outs[False].abs().max().item()
    fl2 = N * (2.0 * V * D + 2 * 2.0 * V2 * D)            # W_q GEMM + Q K^T + P K (the contractions the result needs)
    fl3 = N * (2.0 * V * D + 3 * 2.0 * V2 * D)            # + the statistics pass both paths run
    for fused in (False, True):
        ms = sorted(res[fused])[len(res[fused]) // 2]
        print("%-9s %.3f ms  | %.0f TFLOP/s of executed work (3 passes over the keys), %.0f TFLOP/s of required work (2 passes)"
              % ("fused" if fused else "composed", ms, fl3 / ms / 1e9, fl2 / ms / 1e9))
    print("max |fused - composed| / max |composed| = %.2e" % err)
    # the fused kernel alone
    Q = torch.randn(N, D, device=dev).mul_(0.3).bfloat16()
    row_max = torch.empty(h, N, device=dev); row_inv = torch.empty(h, N, device=dev)
    for i in range(h):
        st = ops.ctc_head_stats(Q[:, i * d:(i + 1) * d], table[:, i * d:(i + 1) * d], None, 1, N, 0, V2, d, 0)
        row_max[i].copy_(st.row_max); torch.reciprocal(st.row_sumexp, out=row_inv[i])
    Z = torch.empty(N, D, device=dev)
    t = []
    for rep in range(4):
        torch.cuda.synchronize()
        e0.record(); ops.attn_softmax_pv(Q, table, N, V2, h, d, Z, row_max, row_inv); e1.record()
        torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    ms = sorted(t[1:])[1]
    print("tasu_attn_softmax_pv, statistics given (one sweep): %.3f ms = %.0f TFLOP/s (2 x 2 N V2 D)" % (ms, N * 2 * 2.0 * V2 * D / ms / 1e9))
    Z2 = torch.empty_like(Z)
    t = []
    for rep in range(4):
        torch.cuda.synchronize()
        e0.record(); ops.attn_softmax_pv(Q, table, N, V2, h, d, Z2); e1.record()
        torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    ms = sorted(t[1:])[1]
    print("tasu_attn_softmax_pv, maxima found in a first sweep: %.3f ms = %.0f TFLOP/s (3 x 2 N V2 D)  max diff %.2e"
          % (ms, N * 3 * 2.0 * V2 * D / ms / 1e9, (Z2 - Z).abs().max().item() / Z.abs().max().item()))
    # A/B of the key split (ops.ATTN_KEY_SPLIT), variants interleaved: the kernel alone and the whole projector
    import ctypes
    import ps_slm_b200._lib as L
    n_splits = ctypes.c_int(0)
    ws_bytes = int(L.lib().tasu_attn_split_plan(N, V2, h, d, ctypes.byref(n_splits)))
    tk, tp = {False: [], True: []}, {False: [], True: []}
    P.FUSED_ATTENTION = True
    for rep in range(5):
        for split in (False, True):
            ops.ATTN_KEY_SPLIT = split
            torch.cuda.synchronize()
            e0.record(); ops.attn_softmax_pv(Q, table, N, V2, h, d, Z2); e1.record()
            torch.cuda.synchronize()
            tk[split].append(e0.elapsed_time(e1))
            e0.record()
            with torch.no_grad():
                ca(post, table)
            e1.record()
            torch.cuda.synchronize()
            tp[split].append(e0.elapsed_time(e1))
    ops.ATTN_KEY_SPLIT = True
    med = lambda v: sorted(v[1:])[len(v[1:]) // 2]
    print("key split A/B (%d items -> %d splits, workspace %.0f MB): kernel %.3f -> %.3f ms, projector %.3f -> %.3f ms"
          % ((N + 127) // 128 * h, n_splits.value, ws_bytes / 1e6, med(tk[False]), med(tk[True]), med(tp[False]), med(tp[True])))
    t = []
    for rep in range(3):
        torch.cuda.synchronize()
        e0.record()
        for i in range(h):
            ops.ctc_head_stats(Q[:, i * d:(i + 1) * d], table[:, i * d:(i + 1) * d], None, 1, N, 0, V2, d, 0)
        e1.record()
        torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    print("statistics pass, 8 heads: %.3f ms = %.0f TFLOP/s" % (min(t), N * 2.0 * V2 * D / min(t) / 1e9))


if __name__ == "__main__":
    main()
