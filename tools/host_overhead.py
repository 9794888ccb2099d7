#!/usr/bin/env python
"""Host-side cost of one TasuBridge call (the launch-bound regime: small batches, 8 ranks on one box): wall time per
call at a tiny batch, and a cProfile of the enqueue path.  `python tools/host_overhead.py [B] [T]` on a GPU box."""
import cProfile
import os
import pstats
import sys
import time
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ps_slm_b200.projector as P
import ps_slm_b200.synth as S
from ps_slm_b200.bridge import TasuBridge


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 83
    dev = torch.device("cuda", 0)
    w, b = S.make_ctc_head()
    torch.manual_seed(0)
    cfg = types.SimpleNamespace(encoder_dim=S.V_CTC, llm_dim=S.H_LLM, encoder_projector_ds_rate=1)
    proj = P.EncoderProjectorLinearSiLU(cfg).to(dev).eval()
    table = S.make_embed_table(dtype=torch.bfloat16, device=dev)
    br = TasuBridge(w.to(dev), b.to(dev), proj, table, S.SPEECH_ID, S.PAD_ID)
    raw, raw_lens, _ = S.make_encoder_batch(B, T, w, seed=3)
    ids, mask, _ = S.make_prompts(B, seed=3, left_pad=True)
    args = tuple(t.to(dev) for t in (raw, raw_lens, ids, mask))
    for _ in range(10):
        br(*args)
    torch.cuda.synchronize()
    n = 200
    t0 = time.perf_counter()
    for _ in range(n):
        br(*args)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print("B=%d T=%d: %.1f us per call (wall, %d calls back to back)" % (B, T, (t1 - t0) / n * 1e6, n))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(n):
        br(*args)
    torch.cuda.synchronize()
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(28)


if __name__ == "__main__":
    main()
