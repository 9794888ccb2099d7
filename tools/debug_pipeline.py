"""Timeline of HostPipeline: per-iteration event timestamps (ms since start) on the three streams."""
import sys, os, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ps_slm_b200.projector as P
import ps_slm_b200.synth as S
from ps_slm_b200.bridge import TasuBridge, HostPipeline

dev = torch.device("cuda:0")
B, T = 64, 500
w, b = S.make_ctc_head()
torch.manual_seed(0)
cfg = types.SimpleNamespace(encoder_dim=S.V_CTC, llm_dim=S.H_LLM, encoder_projector_ds_rate=1)
proj = P.EncoderProjectorLinearSiLU(cfg).to(dev).eval()
table = S.make_embed_table(dtype=torch.bfloat16, device=dev)
bridge = TasuBridge(w.to(dev), b.to(dev), proj, table, S.SPEECH_ID, S.PAD_ID)
host = []
for r in range(4):
    raw, raw_lens, _ = S.make_encoder_batch(B, T, w, seed=r)
    ids, mask, _ = S.make_prompts(B, seed=r, left_pad=True)
    host.append(tuple(t.pin_memory() for t in (raw, raw_lens, ids, mask)))
pipe = HostPipeline(bridge, dev)
for _ in pipe.run(host[i % 4] for i in range(4)):
    pass
torch.cuda.synchronize()

marks = []
orig_upload, orig_download = pipe._upload, pipe._download
def upload(batch):
    e0 = torch.cuda.Event(enable_timing=True); e0.record(pipe.s_in)
    out = orig_upload(batch)
    e1 = torch.cuda.Event(enable_timing=True); e1.record(pipe.s_in)
    marks.append(("h2d", e0, e1)); return out
def download(outs, slot):
    e0 = torch.cuda.Event(enable_timing=True)
    out = orig_download(outs, slot)
    e1 = torch.cuda.Event(enable_timing=True); e1.record(pipe.s_out)
    marks.append(("d2h_end", e1, e1)); return out
pipe._upload, pipe._download = upload, download
orig_call = bridge.__call__
class Wrap:
    def __init__(s, br): s.br = br; s.embed_table = br.embed_table
    def __call__(s, *a):
        e0 = torch.cuda.Event(enable_timing=True); e0.record()
        out = s.br(*a)
        e1 = torch.cuda.Event(enable_timing=True); e1.record()
        marks.append(("comp", e0, e1)); return out
pipe.bridge = Wrap(bridge)
t0 = torch.cuda.Event(enable_timing=True); t0.record(); torch.cuda.synchronize()
import time
w0 = time.perf_counter()
for _ in pipe.run(host[i % 4] for i in range(8)):
    pass
torch.cuda.synchronize()
print("wall ms", (time.perf_counter() - w0) * 1e3)
for name, a, z in marks:
    print(f"{name:8s} {t0.elapsed_time(a):8.3f} -> {t0.elapsed_time(z):8.3f}")
