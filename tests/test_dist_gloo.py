"""CPU, world_size=2 over gloo: utterance sharding, cross-rank packing all-gather and the bucketed
projector-gradient all-reduce (ps-slm_b200/dist.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_utts, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ps_slm_b200.dist as D
    H = 6
    # global ground truth, identical on every rank
    g = torch.Generator().manual_seed(0)
    lens_all = torch.randint(0, 5, (n_utts,), generator=g)
    rows_all = [torch.randn(int(n), H, generator=g) for n in lens_all]
    mine = D.shard_indices(n_utts, rank, world)
    rows = torch.cat([rows_all[i] for i in mine], 0) if mine else torch.zeros(0, H)
    lens = torch.tensor([int(lens_all[i]) for i in mine], dtype=torch.int64)
    rows_g, lens_g = D.all_gather_packed(rows, lens)
    ok = torch.equal(lens_g, lens_all) and torch.equal(rows_g, torch.cat(rows_all, 0))
    # gradient all-reduce (average) over several buckets
    torch.manual_seed(1)
    params = [torch.nn.Parameter(torch.zeros(n)) for n in (7, 1000, 3, 50000)]
    for i, p in enumerate(params):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    D.allreduce_gradients(params, bucket_bytes=4096)
    exp = [(1 + 2) / 2 * (i + 1) for i in range(4)]
    ok = ok and all(torch.allclose(p.grad, torch.full_like(p, e)) for p, e in zip(params, exp))
    # ZeRO-2-style reduce-scatter of the same gradients: this rank's shard of the averaged flat gradient
    for i, p in enumerate(params):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    shard, (lo, hi), n = D.reduce_scatter_gradients(params)
    full = torch.cat([torch.full((p.numel(),), e) for p, e in zip(params, exp)])
    ok = ok and n == full.numel() and torch.allclose(shard[:hi - lo], full[lo:hi]) and (lo, hi) == (rank * shard.numel(), min((rank + 1) * shard.numel(), n))
    # selection straight out of the all-gather buffer (what a rank splices after the length-grouped re-deal)
    g = D.gather_packed(rows, lens, n_global=n_utts)
    sel = [5, 0, 3]
    rows_s, lens_s = g.select(sel)
    ok = ok and torch.equal(lens_s, lens_all[sel]) and torch.equal(rows_s, torch.cat([rows_all[i] for i in sel], 0))
    ok = ok and g.lens_host.tolist() == lens_all.tolist()
    # overlapped all-reduce: first part of one flat gradient buffer reduced early, the rest afterwards
    flat = torch.arange(10.) * (rank + 1)
    ps = [torch.nn.Parameter(torch.zeros(6)), torch.nn.Parameter(torch.zeros(4))]
    D.enable_overlapped_allreduce(True, params=ps)
    hook = D.overlap_hook()
    ok = ok and hook is not None
    hook(flat, 6)
    ps[0].grad, ps[1].grad = flat[:6], flat[6:]
    D.allreduce_gradients(ps)
    ok = ok and torch.allclose(flat, torch.arange(10.) * 1.5)
    # gradient accumulation: the parameters already hold gradients → nothing may be reduced early (a half that is
    # already summed over the ranks would be reduced twice)
    ok = ok and D.overlap_hook() is None
    D.enable_overlapped_allreduce(False)
    out_q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_sharding_is_a_partition():
    import ps_slm_b200.dist as D
    for n in (0, 1, 7, 64):
        for w in (1, 2, 4, 8):
            parts = [D.shard_indices(n, r, w) for r in range(w)]
            assert sorted(sum(parts, [])) == list(range(n))
            order = D.global_order(n, w)
            assert all(parts[r][j] == i for i, (r, j) in enumerate(order))


def test_gather_and_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 7, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_length_grouped_partition():
    import random

    import ps_slm_b200.dist as D
    rnd = random.Random(0)
    for w in (1, 2, 8):
        for n in (0, 1, 5, 64, 513):
            tot = [rnd.choice([25, 40, 110, 160, 230]) + rnd.randint(0, 20) for _ in range(n)]
            parts = D.length_grouped_partition(tot, w)
            assert len(parts) == w and sorted(sum(parts, [])) == list(range(n))
            if n >= 8 * w and w > 1:
                area = [len(p) * max(tot[u] for u in p) for p in parts if p]
                eff = sum(tot) / sum(area)
                naive = [list(range(r, n, w)) for r in range(w)]
                eff_naive = sum(tot) / sum(len(p) * max(tot[u] for u in p) for p in naive)
                assert eff > eff_naive and max(area) <= 1.5 * (sum(area) / len(area)) + 260
