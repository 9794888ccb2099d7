#!/usr/bin/env python
"""Experiment: N independent batches in flight on N CUDA streams (one host thread per stream, one TasuBridge per
stream sharing the weights) vs the single-stream loop of bench.py.  Prints ms per batch for each setting."""
import os
import sys
import threading
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import ps_slm_b200.projector as P
    import ps_slm_b200.synth as S
    from ps_slm_b200.bridge import TasuBridge
    dev = torch.device("cuda", 0)
    B, T, rotate = 64, 500, 4
    w, b = S.make_ctc_head()
    torch.manual_seed(0)
    proj = P.EncoderProjectorLinearSiLU(types.SimpleNamespace(encoder_dim=S.V_CTC, llm_dim=S.H_LLM, encoder_projector_ds_rate=1)).to(dev).eval()
    table = S.make_embed_table(dtype=torch.bfloat16, device=dev)
    wd, bd = w.to(dev), b.to(dev)
    devb = []
    for r in range(rotate):
        raw, raw_lens, _ = S.make_encoder_batch(B, T, w, seed=r)
        ids, mask, _ = S.make_prompts(B, seed=r, left_pad=True)
        devb.append(tuple(t.to(dev) for t in (raw, raw_lens, ids, mask)))
    for n_streams in (1, 2, 3):
        bridges = [TasuBridge(wd, bd, proj, table, S.SPEECH_ID, S.PAD_ID) for _ in range(n_streams)]
        streams = [torch.cuda.Stream(dev) for _ in range(n_streams)]
        steps = 60

        def worker(k, n):
            torch.cuda.set_device(dev)
            with torch.cuda.stream(streams[k]):
                for i in range(n):
                    bridges[k](*devb[(i * n_streams + k) % rotate])

        for phase, n in (("warm", 6), ("timed", steps // n_streams)):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            th = [threading.Thread(target=worker, args=(k, n)) for k in range(n_streams)]
            for t in th:
                t.start()
            for t in th:
                t.join()
            for s in streams:
                torch.cuda.current_stream().wait_stream(s)
            e1.record()
            torch.cuda.synchronize()
            if phase == "timed":
                ms = e0.elapsed_time(e1) / (n * n_streams)
                print(f"streams={n_streams}: {ms:.4f} ms per batch, {B * T / ms * 1e3:.4e} frames/s", flush=True)


if __name__ == "__main__":
    main()
