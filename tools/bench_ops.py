#!/usr/bin/env python
"""Per-operator micro-benchmarks of the drop-in method API (reference-shaped tensors: fp32, odd 25055 pitch):
psd(), the projector module, _merge_input_ids_with_audio_features.  Prints a markdown table."""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import ps_slm_b200.bridge as bridge  # noqa: E402
import ps_slm_b200.ops as ops  # noqa: E402
import ps_slm_b200.projector as P  # noqa: E402
import ps_slm_b200.synth as S  # noqa: E402
import ps_slm_b200._lib as L  # noqa: E402

dev = torch.device("cuda:0")
HBM = 6543.1


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


rows = []
B, T, V = 64, 500, S.V_CTC
w, b = S.make_ctc_head()
raw, raw_lens, _ = S.make_encoder_batch(B, T, w, seed=1)
post = torch.softmax((raw[:, 4:].to(dev) @ w.to(dev).T), -1)          # [B,T,V] fp32, pitch 25055 (rows misaligned)
lens = (raw_lens - 4).to(dev)
ms = timeit(lambda: ops.frame_stats(post, L.INPUT_PROBS, 0))
rows.append(("frame_stats (probs fp32, unpadded pitch)", ms, B * T * V * 4 / ms / 1e6, "GB/s", B * T * V * 4 / ms / 1e6 / HBM))
ms = timeit(lambda: bridge.psd(post, lens, post))
out, nl = bridge.psd(post, lens, post)
n_out = int(nl.sum())
byt = B * T * V * 4 + int(nl.sum()) * V * 4 * 2.3 + B * int(nl.max()) * V * 4
rows.append(("psd() API: stats + plan + sync + padded fp32 mean-pool", ms, byt / ms / 1e6, "GB/s", byt / ms / 1e6 / HBM))
cfg = types.SimpleNamespace(encoder_dim=V, llm_dim=S.H_LLM, encoder_projector_ds_rate=1)
proj = P.EncoderProjectorLinearSiLU(cfg).to(dev).eval()
with torch.no_grad():
    ms = timeit(lambda: proj(out))
fl = out.shape[0] * out.shape[1] * (2.0 * V * 2048 + 2.0 * 2048 * S.H_LLM)
rows.append((f"EncoderProjectorLinearSiLU module on padded [{out.shape[0]},{out.shape[1]},V] fp32", ms, fl / ms / 1e9, "TFLOP/s", fl / ms / 1e9 / 1357.9))
ids, mask, _ = S.make_prompts(B, seed=1, left_pad=True)
ids, mask = ids.to(dev), mask.to(dev)
table = S.make_embed_table(dtype=torch.float32, device=dev)
with torch.no_grad():
    af = proj(out)
emb = torch.nn.functional.embedding(ids, table)
ms = timeit(lambda: bridge.merge_input_ids_with_audio_features(af, nl, emb, ids, mask, None, S.SPEECH_ID, S.PAD_ID))
e = bridge.merge_input_ids_with_audio_features(af, nl, emb, ids, mask, None, S.SPEECH_ID, S.PAD_ID)[0]
byt = (n_out + int(mask.sum()) + e.shape[0] * e.shape[1]) * S.H_LLM * 4
rows.append(("_merge_input_ids_with_audio_features API (fp32)", ms, byt / ms / 1e6, "GB/s", byt / ms / 1e6 / HBM))
# cross-attention projector (projector.py:104-126) on the same compressed batch against the full 151 936-row table
ca = P.EncoderProjectorCTCCA(cfg).to(dev).eval()
table_bf = table.bfloat16()
with torch.no_grad():
    ms = timeit(lambda: ca(out, table_bf), n=3, warm=1)
rows_ca = out.shape[0] * out.shape[1]
fl = rows_ca * (2.0 * V * S.H_LLM + 3 * 2.0 * S.V_LLM * S.H_LLM)          # W_q GEMM + stats pass, softmax pass, P·V
rows.append((f"EncoderProjectorCTCCA (cross-attention over the 151936-row table) on padded [{out.shape[0]},{out.shape[1]},V]", ms,
             fl / ms / 1e9, "TFLOP/s", fl / ms / 1e9 / 1357.9))
# voca_trans branch (ps-slm.py:485-516): simple_linear CTC head over the LLM vocabulary (k = 2), PSD on its logits,
# softmax(no-blank) · embed_matrix — on the raw 512-d encoder frames of the same batch
del ca
torch.cuda.empty_cache()
vcfg = types.SimpleNamespace(encoder_dim=S.D_ENC, llm_dim=151644, encoder_projector_ds_rate=2)
head = P.EncoderProjectorLinear(vcfg).to(dev).eval()
with torch.no_grad():
    head.map.weight.mul_(8.0)
    head.map.bias[151643] = 4.0
enc = raw[:, 4:].to(dev)
with torch.no_grad():
    ms = timeit(lambda: bridge.voca_trans_project(head, enc, lens, table_bf, True, False), n=3, warm=1)
    vo, vl = bridge.voca_trans_project(head, enc, lens, table_bf, True, False)
n_v = B * (T // 2)
fl = n_v * 2.0 * 1024 * 151644 + int(vl.sum()) * 2.0 * 151643 * S.H_LLM
rows.append((f"voca_trans branch: CTC head over 151644 classes on {n_v} concat frames + PSD + softmax·E ({int(vl.sum())} rows kept)", ms,
             fl / ms / 1e9, "TFLOP/s", fl / ms / 1e9 / 1357.9))
print("| op (B=64, T=500) | ms | achieved | unit | frac of measured peak |\n|---|---:|---:|---|---:|")
for r in rows:
    print(f"| {r[0]} | {r[1]:.3f} | {r[2]:.0f} | {r[3]} | {r[4]:.2f} |")
