"""Timeline of the end-to-end host pipeline (bench.py `e2e`): K steps of HostPipeline.run under torch.profiler (CUPTI),
then per stream: busy time, the memcpy durations, and the idle gaps of the compute stream(s) — where the e2e step time
goes beyond the device-resident step.  `python tools/e2e_timeline.py [--bf16] [--steps 12]` on a GPU box."""
import argparse
import json
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bf16", action="store_true")
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--compute-streams", type=int, default=None)
    ap.add_argument("--trace", default=None)
    args = ap.parse_args()
    import ps_slm_b200.projector as P
    import ps_slm_b200.synth as S
    from ps_slm_b200.bridge import HostPipeline, TasuBridge

    dev = torch.device("cuda", 0)
    w, b = S.make_ctc_head()
    torch.manual_seed(0)
    cfg = types.SimpleNamespace(encoder_dim=S.V_CTC, llm_dim=S.H_LLM, encoder_projector_ds_rate=1)
    proj = P.EncoderProjectorLinearSiLU(cfg).to(dev).eval()
    table = S.make_embed_table(dtype=torch.bfloat16, device=dev)
    bridge = TasuBridge(w.to(dev), b.to(dev), proj, table, S.SPEECH_ID, S.PAD_ID)
    host = []
    for r in range(4):
        raw, raw_lens, _ = S.make_encoder_batch(64, 500, w, seed=r)
        if args.bf16:
            raw = raw.bfloat16()
        ids, mask, _ = S.make_prompts(64, seed=r, left_pad=True)
        host.append(tuple(t.pin_memory() for t in (raw, raw_lens, ids, mask)))
    kw = {} if args.compute_streams is None else {"compute_streams": args.compute_streams}
    pipe = HostPipeline(bridge, dev, **kw)

    def run(n):
        for _ in pipe.run(host[i % 4] for i in range(n)):
            pass

    run(8)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(args.steps); e1.record(); torch.cuda.synchronize()
    plain_ms = e0.elapsed_time(e1) / args.steps
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        run(args.steps)
        torch.cuda.synchronize()
    path = args.trace or "/tmp/e2e_trace.json"
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    ev.sort(key=lambda e: e["ts"])
    t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
    streams = {}
    for e in ev:
        streams.setdefault(e["args"].get("stream"), []).append(e)
    rep = {"ms_per_step_unprofiled": plain_ms, "ms_per_step_profiled": (t1 - t0) / 1e3 / args.steps, "streams": {}}
    for s, es in streams.items():
        busy = sum(e["dur"] for e in es)
        kinds = {}
        for e in es:
            nm = e["name"][:48]
            k = kinds.setdefault(nm, [0, 0.0])
            k[0] += 1
            k[1] += e["dur"]
        gaps = []
        for a, z in zip(es, es[1:]):
            g = z["ts"] - (a["ts"] + a["dur"])
            if g > 20:
                gaps.append((round(g), a["name"][:32], z["name"][:32]))
        gaps.sort(reverse=True)
        rep["streams"][str(s)] = {
            "busy_ms_per_step": busy / 1e3 / args.steps, "events": len(es),
            "top": sorted(([n, c, round(d / 1e3 / args.steps, 4)] for n, (c, d) in kinds.items()), key=lambda x: -x[2])[:8],
            "gap_ms_per_step_over_20us": sum(g[0] for g in gaps) / 1e3 / args.steps, "largest_gaps_us": gaps[:10]}
    print(json.dumps(rep, indent=1))


if __name__ == "__main__":
    main()
