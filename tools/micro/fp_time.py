"""Back-to-back timing of tasu_fingerprint (queue kept full: elapsed / launches = kernel time)."""
import sys, types, torch
sys.path.insert(0, "/root/repo")
import ps_slm_b200.ops as ops, ps_slm_b200.synth as S, ps_slm_b200.projector as P
dev = torch.device("cuda:0")
w, b = S.make_ctc_head()
proj = P.EncoderProjectorLinearSiLU(types.SimpleNamespace(encoder_dim=S.V_CTC, llm_dim=S.H_LLM, encoder_projector_ds_rate=1)).to(dev)
params = [w.to(dev), b.to(dev)] + list(proj.parameters())[:6]
out_dev = torch.zeros(1, dtype=torch.int64, device=dev)
out_pin = torch.zeros(1, dtype=torch.int64).pin_memory()
big = torch.empty(1 << 28, dtype=torch.uint8, device=dev)
for name, out, ps in (("8 buffers, device out", out_dev, params), ("8 buffers, pinned host out", out_pin, params),
                      ("1 buffer (W1), device out", out_dev, params[4:5]), ("1 small buffer, device out", out_dev, params[1:2])):
    for _ in range(10):
        ops.fingerprint(ps, out=out)
    torch.cuda.synchronize()
    n = 300
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    big.zero_(); big.zero_(); big.zero_()                     # ~0.2 ms of queued work: the launches below pile up behind it
    e0.record()
    for _ in range(n):
        ops.fingerprint(ps, out=out)
    e1.record(); torch.cuda.synchronize()
    print("%-28s %.2f us per launch" % (name, e0.elapsed_time(e1) * 1e3 / n))
