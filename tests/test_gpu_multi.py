"""Multi-GPU (NCCL) test of BASELINE configs[3]: a mixed multitask batch is utterance-sharded over the
ranks, every rank compresses + projects its shard, lengths and packed rows are all-gathered, and the
globally packed splice must equal the single-GPU result bit for bit.  Needs >= 2 GPUs (skipped otherwise;
the host-side logic is covered on CPU by tests/test_dist_gloo.py)."""
import os
import socket
import types

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import ps_slm_b200.dist as D
    import ps_slm_b200.projector as P
    import ps_slm_b200.synth as S
    from ps_slm_b200.bridge import TasuBridge
    B, T = 10, 90
    w, b = S.make_ctc_head()
    torch.manual_seed(0)
    proj = P.EncoderProjectorLinearSiLU(types.SimpleNamespace(encoder_dim=S.V_CTC, llm_dim=S.H_LLM, encoder_projector_ds_rate=1)).to(dev).eval()
    table = S.make_embed_table(dtype=torch.bfloat16, device=dev)
    br = TasuBridge(w.to(dev), b.to(dev), proj, table, S.SPEECH_ID, S.PAD_ID)
    raw, raw_lens, _ = S.make_encoder_batch(B, T, w, seed=5, ragged=True)
    tasks = ["ASR", "EN2ZH", "EN2DE", "QA", "SLU_scenario"]
    ids, mask, _ = S.make_prompts(B, seed=5, tasks=tasks, left_pad=True)
    mine = D.shard_indices(B, rank, world)
    audio, lens, _ = br.compress_project(raw[mine].to(dev), raw_lens[mine].to(dev))
    rows_g, lens_g = D.all_gather_packed(audio, lens, n_global=B)
    out = br.splice(rows_g, lens_g, ids.to(dev), mask.to(dev))
    # single-GPU reference on the same device
    ref = br(raw.to(dev), raw_lens.to(dev), ids.to(dev), mask.to(dev))
    ok = (torch.equal(lens_g, ref[4]) and torch.equal(out[0], ref[0]) and torch.equal(out[1], ref[1])
          and torch.equal(out[3], ref[3]))
    valid = int(out[1].sum())
    # the whole packed step: length-grouped re-deal, this rank's utterances picked out of the all-gather buffer
    plen = mask.sum(1).tolist()
    outs, info = D.packed_inference_step(br, raw[mine].to(dev), raw_lens[mine].to(dev), ids.to(dev), mask.to(dev), plen)
    sels = [None] * world
    dist.all_gather_object(sels, info["sel"])
    ok = ok and sorted(sum(sels, [])) == list(range(B))            # the re-deal is a partition of the global batch
    if outs is not None:
        emb_s, mask_s, _, pos_s, _ = outs
        for k, u in enumerate(info["sel"]):                        # left-padded: compare the valid tail of every row
            n = int(ref[1][u].sum())
            ok = ok and int(mask_s[k].sum()) == n and torch.equal(emb_s[k, emb_s.shape[1] - n:], ref[0][u, ref[0].shape[1] - n:])
            ok = ok and torch.equal(pos_s[k, pos_s.shape[1] - n:], ref[3][u, ref[3].shape[1] - n:])
    q.put((rank, bool(ok), valid / out[1].numel()))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_cross_rank_packing_matches_single_gpu():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert [r[:2] for r in res] == [(0, True), (1, True)], res
    assert 0.3 < res[0][2] <= 1.0          # packing efficiency (valid / padded tokens) is reported


def _worker_overlap(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import ps_slm_b200.dist as D
    import ps_slm_b200.ops as ops
    import ps_slm_b200.projector as P
    import ps_slm_b200.synth as S
    V, H = S.V_CTC, S.H_LLM
    torch.manual_seed(0)
    proj = P.EncoderProjectorLinearSiLU(types.SimpleNamespace(encoder_dim=V, llm_dim=H, encoder_projector_ds_rate=1)).to(dev).train()
    ids = S.make_transcripts(12, V, seed=3, lo=5, hi=30)
    mine = D.shard_indices(12, rank, world)
    torch.manual_seed(100 + rank)
    rows = ops.sim_token_rows(ops.TokenBatch([ids[i] for i in mine]), V, dev)
    gy = torch.randn(rows.n_rows, H, device=dev, generator=torch.Generator(device=dev).manual_seed(rank))
    params = list(proj.parameters())

    def grads(overlap):
        D.enable_overlapped_allreduce(overlap, params=params)
        for p in params:
            p.grad = None
        y = proj.forward_token_rows(rows)
        (y * gy).sum().backward()
        if overlap:
            D.allreduce_gradients(params)                       # W1 half already in flight from inside backward
            return [p.grad.clone() for p in params]
        out = []
        for p in params:                                        # reference: plain per-tensor all-reduce, then average
            g = p.grad.clone()
            dist.all_reduce(g)
            out.append(g / world)
        return out
    ref = grads(False)
    got = grads(True)
    D.enable_overlapped_allreduce(False)
    ok = all(torch.allclose(a, b, rtol=1e-6, atol=1e-9) for a, b in zip(ref, got))
    # bf16 on the wire: one message of half the size, averaged back into the fp32 gradients
    for p in params:
        p.grad = None
    y = proj.forward_token_rows(rows)
    (y * gy).sum().backward()
    D.allreduce_gradients(params, wire_dtype=torch.bfloat16)
    for a, p in zip(ref, params):
        scale = a.abs().max().item() + 1e-12
        ok = ok and ((p.grad - a).abs().max().item() / scale) < 2e-2
    # the default exchange (ONE in-place message, averaged inside NCCL) and the ZeRO-2 reduce-scatter of the same flat
    # buffer (conf/ds_config.json:15-21): this rank's shard of the averaged gradient
    for p in params:
        p.grad = None
    y = proj.forward_token_rows(rows)
    (y * gy).sum().backward()
    shard, (lo, hi), n = D.reduce_scatter_gradients(params)
    D.allreduce_gradients(params)
    ok = ok and all(torch.allclose(p.grad, a, rtol=1e-5, atol=1e-8) for a, p in zip(ref, params))
    flat = D._shared_flat([p.grad for p in params])
    ok = ok and flat is not None and n == flat.numel() and hi - lo == shard.numel() == n // world
    ok = ok and torch.allclose(shard, flat[lo:hi], rtol=1e-5, atol=1e-8)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_overlapped_gradient_allreduce_matches_plain():
    """The all-reduce started inside the token-row backward (W1 half under the W2 half) gives the same averaged
    gradients as reducing every tensor after the backward."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_overlap, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)], res
