#!/usr/bin/env python
"""BASELINE.json configs[4]: throughput sweep, utterance length 5-60 s x batch 16-1024 on the
compress+project(+splice) bridge, one B200 (run under torchrun for the per-GPU sharded variant).
Writes a markdown table + JSON to gpurun_out/sweep.{md,json}."""
import argparse
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, nargs="*", default=[83, 167, 250, 500, 1000])
    ap.add_argument("--batches", type=int, nargs="*", default=[16, 64, 256, 1024])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--cpu-frames-per-s", type=float, default=0.0, help="reference CPU path figure to compare with")
    args = ap.parse_args()
    import torch

    import ps_slm_b200.projector as P
    import ps_slm_b200.synth as S
    from ps_slm_b200.bridge import TasuBridge
    import torch.distributed as dist
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:                                       # weak scaling: every rank runs the same per-GPU batch size
        dist.init_process_group("nccl", device_id=dev)
    w, b = S.make_ctc_head()
    torch.manual_seed(0)
    proj = P.EncoderProjectorLinearSiLU(types.SimpleNamespace(encoder_dim=S.V_CTC, llm_dim=S.H_LLM, encoder_projector_ds_rate=1)).to(dev).eval()
    table = S.make_embed_table(dtype=torch.bfloat16, device=dev)
    bridge = TasuBridge(w.to(dev), b.to(dev), proj, table, S.SPEECH_ID, S.PAD_ID)
    res = []
    for T in args.frames:
        for B in args.batches:
            # build the batch from a 64-utterance seed batch tiled to B (generation on the host is the slow part)
            base_b = min(B, 64)
            raw, raw_lens, _ = S.make_encoder_batch(base_b, T, w, seed=T + 7919 * rank)
            ids, mask, _ = S.make_prompts(base_b, seed=T, left_pad=True)
            rep = B // base_b
            raw, raw_lens, ids, mask = (t.to(dev).repeat(rep, *([1] * (t.dim() - 1))) for t in (raw, raw_lens, ids, mask))
            for _ in range(3):
                out = bridge(raw, raw_lens, ids, mask)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                out = bridge(raw, raw_lens, ids, mask)
            e1.record()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            c = bridge.last_counts
            n_out = c["n_out"]
            if world > 1:                               # device time = max over ranks; rows summed over ranks
                t = torch.tensor([ms], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t)
                t = torch.tensor([n_out], dtype=torch.int64, device=dev)
                dist.all_reduce(t)
                n_out = int(t)
            res.append({"seconds": round(T * 0.06, 1), "T": T, "B": B, "n_gpus": world, "ms_per_step": ms,
                        "frames_in_per_s": world * B * T / (ms / 1e3), "frames_out_per_s": n_out / (ms / 1e3),
                        "compression": world * B * T / max(n_out, 1), "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9})
            del out, raw
            torch.cuda.empty_cache()
            torch.cuda.reset_peak_memory_stats()
            if rank == 0:
                print(res[-1], flush=True)
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "sweep%s.json" % ("" if world == 1 else "_n%d" % world)), "w") as f:
        json.dump(res, f, indent=1)
    with open(os.path.join(ROOT, "gpurun_out", "sweep%s.md" % ("" if world == 1 else "_n%d" % world)), "w") as f:
        f.write("| utt length | batch per GPU | ms/step | frames in /s | rows out /s | compression | peak mem GB |"
                + (" vs CPU ref |" if args.cpu_frames_per_s else "") + "\n|---|---:|---:|---:|---:|---:|---:|"
                + ("---:|" if args.cpu_frames_per_s else "") + "\n")
        for r in res:
            f.write(f"| {r['seconds']} s | {r['B']} | {r['ms_per_step']:.3f} | {r['frames_in_per_s']:.3e} | {r['frames_out_per_s']:.3e} | "
                    f"{r['compression']:.2f} | {r['peak_mem_gb']:.1f} |"
                    + (f" {r['frames_in_per_s'] / args.cpu_frames_per_s:.0f}x |" if args.cpu_frames_per_s else "") + "\n")


if __name__ == "__main__":
    main()
