#!/usr/bin/env python
"""BASELINE.json configs[3]: mixed audio/text multitask batch with variable-length compressed sequences
packed across the GPUs of one box.

    python tools/bench_mixed.py                                               # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 tools/bench_mixed.py    # utterance-sharded over 8 GPUs

Per step and rank (64 utterances per rank, tasks uniform over ASR / EN2ZH / EN2DE / QA / SLU_scenario prompts,
utterance lengths U(5 s, 30 s), half audio, half text-simulated):
  audio half : ctc_lo + stats -> collapse -> kept-frame softmax GEMM -> pooling -> projector   (TasuBridge.compress_project)
  text half  : simulator descriptors -> token-row projector                                    (forward_token_rows)
  exchange   : all-gather of the compressed lengths and of the packed [sum M_b, 1536] bf16 rows (dist.all_gather_packed)
  pack       : utterances are re-dealt to the ranks in contiguous groups of the length-sorted batch with balanced
               padded areas (dist.length_grouped_partition); each rank splices its group (TasuBridge.splice)
Reports frames/s (audio encoder frames + text token rows consumed), the all-gather time and bus bytes, and the
packing efficiency (valid / padded tokens) of the naive per-shard batches vs the length-balanced re-deal.
"""
import argparse
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch-per-gpu", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist

    import ps_slm_b200.dist as D
    import ps_slm_b200.ops as ops
    import ps_slm_b200.projector as P
    import ps_slm_b200.synth as S
    from ps_slm_b200.bridge import TasuBridge

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    V, H, T = S.V_CTC, S.H_LLM, 500
    Bl = args.batch_per_gpu
    Bg = Bl * world
    w, b = S.make_ctc_head()
    torch.manual_seed(0)
    proj = P.EncoderProjectorLinearSiLU(types.SimpleNamespace(encoder_dim=V, llm_dim=H, encoder_projector_ds_rate=1)).to(dev).eval()
    table = S.make_embed_table(dtype=torch.bfloat16, device=dev)
    br = TasuBridge(w.to(dev), b.to(dev), proj, table, S.SPEECH_ID, S.PAD_ID)
    tasks = ["ASR", "EN2ZH", "EN2DE", "QA", "SLU_scenario"]
    ids_g, mask_g, _ = S.make_prompts(Bg, seed=4, tasks=tasks, left_pad=True)          # identical on every rank
    prompt_len = mask_g.sum(1).tolist()
    mine = D.shard_indices(Bg, rank, world)
    # global utterance i is AUDIO when (i // world) is even, TEXT otherwise: a 50/50 mix on every rank
    is_audio = [(i // world) % 2 == 0 for i in mine]
    a_loc = [j for j, a in enumerate(is_audio) if a]
    t_loc = [j for j, a in enumerate(is_audio) if not a]
    g = torch.Generator().manual_seed(100 + rank)
    n_batches = 4
    feeds = []
    for k in range(n_batches):
        raw, raw_lens, _ = S.make_encoder_batch(len(a_loc), T, w, seed=1000 * rank + k)
        frames = torch.randint(83, T + 1, (len(a_loc),), generator=g)                   # 5 s .. 30 s at 60 ms
        raw_lens = frames + S.N_PREFIX
        secs = torch.randint(5, 31, (len(t_loc),), generator=g)
        trans = [S.make_transcripts(1, V, seed=int(7 * rank + 13 * k + j), lo=max(1, int(3.5 * s) - 2), hi=int(3.5 * s) + 2)[0]
                 for j, s in enumerate(secs.tolist())]
        feeds.append((raw.to(dev), raw_lens.to(dev), ops.TokenBatch(trans), int(frames.sum()), sum(len(t) for t in trans)))
    ids_g, mask_g = ids_g.to(dev), mask_g.to(dev)
    perm_local = torch.tensor(a_loc + t_loc)                                             # rows come out audio-first
    inv_local = torch.argsort(perm_local).to(dev)
    src_of_local = np.argsort(np.asarray(a_loc + t_loc))                                 # position of local utterance j in that order
    timing = []
    stats = {}

    def step(i):
        raw, raw_lens, tb, n_frames, n_tok = feeds[i % n_batches]
        torch.manual_seed(50 + i)
        rows_a, lens_a, _ = br.compress_project(raw, raw_lens)
        tr = ops.sim_token_rows(tb, V, dev)
        with torch.no_grad():
            rows_t = proj.forward_token_rows(tr, torch.bfloat16)
        # local utterance order: lengths permuted back to shard order, rows re-ordered by one row gather
        lens_cat = torch.cat([lens_a, tr.lens])
        lens_loc = lens_cat[inv_local]
        lens_host = lens_cat.cpu().numpy()
        offs = np.concatenate([[0], np.cumsum(lens_host)])
        idx = torch.from_numpy(D.concat_ranges(offs[:-1][src_of_local], lens_host[src_of_local]))
        rows_loc = D._gather_rows(torch.cat([rows_a, rows_t]), idx.pin_memory().to(dev, non_blocking=True))
        rows_g, lens_gd, lens_gh = D.all_gather_packed(rows_loc, lens_loc, timing=timing, return_host=True)
        # length-grouped re-deal: contiguous groups of the length-sorted batch with balanced padded areas
        tot = [prompt_len[u] + int(lens_gh[u]) - 1 for u in range(Bg)]
        share = D.length_grouped_partition(tot, world)
        sel = sorted(share[rank])
        lens_gn = np.asarray(lens_gh, dtype=np.int64)
        goff = np.concatenate([[0], np.cumsum(lens_gn)])
        sel_n = np.asarray(sel, dtype=np.int64)
        idx2 = D.concat_ranges(goff[:-1][sel_n], lens_gn[sel_n])
        rows_sel = D._gather_rows(rows_g, torch.from_numpy(idx2).pin_memory().to(dev, non_blocking=True))
        sel_t = torch.tensor(sel, device=dev)
        ids_s, mask_s = ids_g[sel_t], mask_g[sel_t]
        if sel:
            cut = int(ids_s.shape[1] - max(prompt_len[u] for u in sel))                  # drop all-pad columns
            emb, m, _, _, _ = br.splice(rows_sel, lens_gd[sel_t], ids_s[:, cut:].contiguous(), mask_s[:, cut:].contiguous())
        else:
            emb = torch.empty(0, 0, H, dtype=table.dtype, device=dev)
        if i == 0 or "eff_packed" not in stats:
            naive_valid = sum(tot[u] for u in mine)
            naive_pad = len(mine) * max(tot[u] for u in mine)
            stats.update(eff_naive=(naive_valid, naive_pad), eff_packed=(sum(tot[u] for u in sel), emb.shape[0] * emb.shape[1]),
                         spliced_len=emb.shape[1])
        return n_frames + n_tok, int(rows_g.shape[0])

    for i in range(args.warmup):
        step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    timing.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    units = 0
    n_out = 0
    for i in range(args.steps):
        u, n_out = step(args.warmup + i)
        units += u
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ag_ms = sum(a.elapsed_time(b_) for a, b_, _ in timing) / max(len(timing), 1)
    ag_bytes = timing[0][2] if timing else 0
    vals = torch.tensor([ms, ag_ms, float(units), *map(float, stats["eff_naive"]), *map(float, stats["eff_packed"])],
                        dtype=torch.float64, device=dev)
    if world > 1:
        mx = vals.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone()
        dist.all_reduce(sm)
        ms, ag_ms = float(mx[0]), float(mx[1])
        units = float(sm[2])
        naive = (float(sm[3]), float(sm[4]))
        packed = (float(sm[5]), float(sm[6]))
    else:
        naive, packed = stats["eff_naive"], stats["eff_packed"]
    if rank == 0:
        per_step = ms / args.steps
        print(json.dumps({
            "workload": "configs[3] mixed audio/text multitask batch, compressed sequences packed across the box",
            "n_gpus": world, "batch_per_gpu": Bl, "global_batch": Bg, "ms_per_step": per_step,
            "frames_per_s": units / (ms / 1e3), "compressed_rows_global": n_out,
            "allgather_ms": ag_ms, "allgather_bytes_per_rank": ag_bytes,
            "allgather_busbw_gbs": (ag_bytes * (world - 1) / world) / (ag_ms / 1e3) / 1e9 if ag_ms > 0 else None,
            "packing_efficiency_naive_shards": naive[0] / naive[1], "packing_efficiency_length_grouped": packed[0] / packed[1],
            "spliced_len_rank0": stats["spliced_len"]}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
