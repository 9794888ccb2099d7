"""Kernel-by-kernel time of the voca_trans branch (bridge.voca_trans_project) at the config-2 size, fused formulation
(no logits tensor) against the materialised one, via torch.profiler (CUPTI).  `python tools/prof_voca.py` on a GPU box."""
import os, sys, types
sys.path.insert(0, os.getcwd())
import torch
import ps_slm_b200.bridge as bridge, ps_slm_b200.projector as P, ps_slm_b200.synth as S
dev = torch.device("cuda:0")
B, T = 64, 500
w, b = S.make_ctc_head()
raw, raw_lens, _ = S.make_encoder_batch(B, T, w, seed=1)
lens = (raw_lens - 4).to(dev)
table_bf = S.make_embed_table(dtype=torch.bfloat16, device=dev)
vcfg = types.SimpleNamespace(encoder_dim=S.D_ENC, llm_dim=151644, encoder_projector_ds_rate=2)
head = P.EncoderProjectorLinear(vcfg).to(dev).eval()
with torch.no_grad():
    head.map.weight.mul_(8.0); head.map.bias[151643] = 4.0
enc = raw[:, 4:].to(dev)
from torch.profiler import profile, ProfilerActivity
for mode in (True, False):
    bridge.FUSED_VOCA_TRANS = mode
    with torch.no_grad():
        for _ in range(2):
            bridge.voca_trans_project(head, enc, lens, table_bf, True, False)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            bridge.voca_trans_project(head, enc, lens, table_bf, True, False)
            torch.cuda.synchronize()
    print("==== fused" if mode else "==== materialised")
    ev = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:12]
    for e in ev:
        print("%9.1f us x%d  %s" % (e.device_time_total, e.count, e.key[:90]))
    print("total %.1f us" % sum(e.device_time_total for e in prof.key_averages()))
