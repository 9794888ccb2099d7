"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink on the box,
gloo in CPU tests).  The bridge shards by utterance with no data-path collective; NCCL is used
only to all-gather per-rank compressed lengths + packed outputs for cross-rank packing and to
all-reduce the projector gradients in training.

* utterance sharding follows the reference's sample sharding ``i % world == rank``
  (Multitask/dataset/speech_dataset_large.py:80-91);
* the gradient all-reduce replaces DeepSpeed ZeRO-2's reduce-scatter of the 54.5 M trainable
  projector parameters (Multitask/conf/ds_config.json:15-21, finetune_deepspeed.py:147-149).
"""
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def bind_to_local_numa(device_index: int) -> Optional[List[int]]:
    """Pin this process to the CPUs NVML reports as local to GPU ``device_index`` (its NUMA node), so that pinned
    staging buffers allocated afterwards are first-touched on the memory next to the GPU's PCIe root.  With one
    process per GPU this keeps 8 ranks from pulling all their host↔device traffic through one socket.  Returns the
    CPU list, or None when NVML / affinity is unavailable (nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[device_index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else device_index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:  # noqa: BLE001 — affinity is an optimisation, never a requirement
        return None


def bind_rank_to_cpu_slice(local_rank: int, local_world: int) -> Optional[List[int]]:
    """One process per GPU on one box: give every rank its own contiguous slice of the CPUs this process may use
    (rank r gets CPUs [r k, r k + k), k = n_cpus // ranks), so that the ranks' launch threads, NCCL proxies and host
    simulators do not migrate onto each other's cores.  The bridge's steps are host-latency sensitive (one header
    hand-off per batch, ~40 launches per training step); on a 32-vCPU box with 8 ranks an unpinned run showed a
    1.03 -> 1.51 ms training step without any collective (profiles/r02p_bench_n8.json).  Returns the CPU list, or None
    when there are fewer than 2 CPUs per rank / affinity is unavailable (nothing is changed then)."""
    import os
    try:
        cpus = sorted(os.sched_getaffinity(0))
        k = len(cpus) // max(local_world, 1)
        if local_world < 2 or k < 2:
            return None
        mine = cpus[local_rank * k:(local_rank + 1) * k]
        os.sched_setaffinity(0, mine)
        return mine
    except Exception:  # noqa: BLE001 — affinity is an optimisation, never a requirement
        return None


def shard_indices(n: int, rank: int, world_size: int) -> List[int]:
    """Global utterance indices owned by ``rank`` (utterance i → rank i % W)."""
    return list(range(rank, n, world_size))


def global_order(n: int, world_size: int) -> List[Tuple[int, int]]:
    """For every global utterance i: (owner rank, local index)."""
    return [(i % world_size, i // world_size) for i in range(n)]


def length_grouped_partition(total_lens: Sequence[int], world_size: int) -> List[List[int]]:
    """Re-deal a global batch to ``world_size`` ranks for padded batching: utterances are sorted by spliced length and
    cut into contiguous groups (so every rank pads to a length close to its own sequences) whose PADDED areas
    ``count x max_len`` are balanced — the smallest area cap for which a greedy cut needs at most ``world_size``
    groups (binary search).  Returns the utterance indices per rank (ranks may hold different counts)."""
    n = len(total_lens)
    order = sorted(range(n), key=lambda u: (-int(total_lens[u]), u))
    if n == 0:
        return [[] for _ in range(world_size)]

    def cut(cap):
        groups, i = [], 0
        while i < n:
            mx = max(int(total_lens[order[i]]), 1)
            k = max(1, cap // mx)
            groups.append(order[i:i + k])
            i += k
        return groups

    lo, hi = max(int(total_lens[order[0]]), 1), max(int(total_lens[order[0]]), 1) * n
    while lo < hi:
        mid = (lo + hi) // 2
        if len(cut(mid)) <= world_size:
            hi = mid
        else:
            lo = mid + 1
    groups = cut(lo)
    return groups + [[] for _ in range(world_size - len(groups))]


def all_gather_lengths(lens: torch.Tensor, group=None) -> List[torch.Tensor]:
    """All-gather of the per-rank compressed lengths (ragged: ranks may own different counts)."""
    rank, W = world()
    if W == 1:
        return [lens]
    n = torch.tensor([lens.numel()], dtype=torch.int64, device=lens.device)
    counts = [torch.zeros_like(n) for _ in range(W)]
    dist.all_gather(counts, n, group=group)
    m = int(max(int(c) for c in counts))
    pad = torch.zeros(m, dtype=lens.dtype, device=lens.device)
    pad[:lens.numel()] = lens
    bufs = [torch.zeros_like(pad) for _ in range(W)]
    dist.all_gather(bufs, pad, group=group)
    return [b[:int(c)] for b, c in zip(bufs, counts)]


def concat_ranges(starts, lens):
    """int32 index vector of the concatenated ranges [starts[i], starts[i] + lens[i]) — vectorised (no Python loop)."""
    import numpy as np
    starts = np.asarray(starts, dtype=np.int64)
    lens = np.asarray(lens, dtype=np.int64)
    total = int(lens.sum())
    if total == 0:
        return np.zeros(0, dtype=np.int32)
    before = np.concatenate([[0], np.cumsum(lens)[:-1]])
    return (np.repeat(starts - before, lens) + np.arange(total)).astype(np.int32)


def _gather_rows(flat: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """rows[i] = flat[idx[i]] — tasu_gather_rows on the GPU; plain indexing only for the CPU (gloo) tests."""
    if flat.is_cuda:
        from . import _lib as L
        from . import ops
        H = flat.shape[1]
        out = torch.empty(max(idx.numel(), 1), H, dtype=flat.dtype, device=flat.device)[:idx.numel()]
        L.check(L.lib().tasu_gather_rows(flat.data_ptr(), ops._dt(flat), flat.stride(0), idx.data_ptr(), idx.numel(), H,
                                         out.data_ptr(), H, ops._stream()), "tasu_gather_rows")
        ops._count(1)
        return out
    return flat.index_select(0, idx.long())


_PINNED = {}


def _pinned_i64(n: int, slot: int = 0) -> torch.Tensor:
    """Small ring of reusable pinned int64 host buffers (pinning memory per call costs more than the exchange itself)."""
    key = (slot, max(64, 1 << (max(n, 1) - 1).bit_length()))
    buf = _PINNED.get(key)
    if buf is None:
        buf = torch.empty(key[1], dtype=torch.int64, pin_memory=torch.cuda.is_available())
        _PINNED[key] = buf
    return buf[:n]


class PackedGather:
    """Result of ``gather_packed``: the flat all-gather buffer of every rank's packed rows (rank r's rows start at row
    ``r * slab_rows``), the gathered lengths on the device (``all_lens [W, b_max]`` int64) and on the host
    (``lens_host``: numpy [n_global] in GLOBAL utterance order, utterance i = rank i % W, local index i // W)."""
    __slots__ = ("flat", "all_lens", "lens_host", "slab_rows", "W", "b_max", "n_global")

    def select(self, sel: Optional[Sequence[int]] = None):
        """Rows and lengths of the utterances ``sel`` (global ids, in that order; None = all, global order):
        ``(rows [sum len, H], lens int64 [n_sel])`` — ``tasu_packed_select`` copies them straight out of the
        all-gather buffer (2 launches; only the ≤ n_sel utterance ids travel to the device)."""
        import numpy as np
        n_sel = self.n_global if sel is None else len(sel)
        H = self.flat.shape[1]
        dev = self.flat.device
        sel_np = np.arange(self.n_global, dtype=np.int64) if sel is None else np.asarray(sel, dtype=np.int64)
        lens_sel = self.lens_host[sel_np] if n_sel else np.zeros(0, dtype=np.int64)
        total = int(lens_sel.sum())
        if not self.flat.is_cuda:                                # CPU (gloo) tests: plain indexing
            r_of, j_of = sel_np % self.W, sel_np // self.W
            local_off = np.zeros((self.W, self.b_max + 1), dtype=np.int64)
            lens2 = np.zeros((self.W, self.b_max), dtype=np.int64)
            gi = np.arange(self.n_global)
            lens2[gi % self.W, gi // self.W] = self.lens_host
            local_off[:, 1:] = np.cumsum(lens2, axis=1)
            starts = r_of * self.slab_rows + local_off[r_of, j_of]
            idx = torch.from_numpy(concat_ranges(starts, lens_sel))
            return _gather_rows(self.flat, idx), torch.from_numpy(lens_sel.copy())
        from . import _lib as L
        from . import ops
        out = torch.empty(max(total, 1), H, dtype=self.flat.dtype, device=dev)[:total]
        out_lens = torch.empty(max(n_sel, 1), dtype=torch.int64, device=dev)[:n_sel]
        ws = torch.empty(self.W * self.b_max + 2 * n_sel + 1, dtype=torch.int32, device=dev)
        sel_dev = None
        if sel is not None and n_sel:
            host = _pinned_i64(n_sel, slot=1)
            host.copy_(torch.from_numpy(sel_np))
            sel_dev = host.to(dev, non_blocking=True).to(torch.int32)
        L.check(L.lib().tasu_packed_select(self.flat.data_ptr(), ops._dt(self.flat), self.flat.stride(0), self.slab_rows, H,
                                           self.all_lens.data_ptr(), self.W, self.b_max, self.n_global,
                                           sel_dev.data_ptr() if sel_dev is not None else None, n_sel, out.data_ptr(), H,
                                           total, out_lens.data_ptr(), ws[-1:].data_ptr(), ws.data_ptr(), ops._stream()),
                "tasu_packed_select")
        ops._count(2)
        return out, out_lens


def gather_packed(rows: torch.Tensor, lens: torch.Tensor, n_global: Optional[int] = None, group=None,
                  timing=None) -> PackedGather:
    """All-gather every rank's packed compressed rows ``[sum M_b, H]`` and lengths (north star: the only data-path
    collectives of inference).  One fixed-shape all-gather of the lengths, ONE device→host read of them (pinned,
    reused buffer), one all-gather of the payload padded to the largest rank's row count (on NVSwitch a flat all-gather
    is bandwidth-optimal).  ``n_global``: number of utterances of the global batch when known (utterance i lives on
    rank i % W); without it the per-rank counts are exchanged first.  ``timing``: list that receives
    ``(start event, end event, bytes received per rank)`` of the payload all-gather."""
    import numpy as np
    rank, W = world()
    n_local = lens.numel()
    if n_global is None:
        if W == 1:
            n_global = n_local
        else:
            cnt = torch.tensor([n_local], dtype=torch.int64, device=lens.device)
            counts = [torch.zeros_like(cnt) for _ in range(W)]
            dist.all_gather(counts, cnt, group=group)
            n_global = int(sum(int(c) for c in counts))
    b_max = (n_global + W - 1) // W
    if n_local != len(shard_indices(n_global, rank, W)):
        raise ValueError("rank %d holds %d utterances, the i %% W sharding of %d gives %d"
                         % (rank, n_local, n_global, len(shard_indices(n_global, rank, W))))
    dev = rows.device
    H = rows.shape[1]
    pad_lens = torch.zeros(max(b_max, 1), dtype=torch.int64, device=dev)
    pad_lens[:n_local] = lens.to(torch.int64)
    all_lens = torch.empty(W, max(b_max, 1), dtype=torch.int64, device=dev)
    if W > 1:
        dist.all_gather_into_tensor(all_lens.view(-1), pad_lens, group=group)
    else:
        all_lens[0] = pad_lens
    # the single device→host hand-off of the exchange
    if dev.type == "cuda":
        host = _pinned_i64(all_lens.numel(), slot=0)
        host.copy_(all_lens.view(-1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        lens2 = host.numpy().reshape(W, max(b_max, 1)).copy()
    else:
        lens2 = all_lens.numpy().copy()
    gi = np.arange(n_global)
    g = PackedGather()
    g.W, g.b_max, g.n_global = W, max(b_max, 1), n_global
    g.lens_host = lens2[gi % W, gi // W] if n_global else np.zeros(0, dtype=np.int64)
    totals = lens2.sum(axis=1)
    if int(totals[rank]) != rows.shape[0]:
        raise ValueError("rank %d: %d packed rows but the lengths sum to %d" % (rank, rows.shape[0], int(totals[rank])))
    m = int(max(int(totals.max()), 1))
    g.slab_rows, g.all_lens = m, all_lens
    if W == 1:
        g.flat = rows
        return g
    # payload, padded to the largest rank: a view when the caller's buffer is large enough, else one copy
    base = rows._base if rows._base is not None else rows
    if (rows.is_contiguous() and base.dim() == 2 and base.shape[1] == H and base.is_contiguous()
            and base.data_ptr() == rows.data_ptr() and base.shape[0] >= m):
        pad = base[:m]
    else:
        pad = torch.empty(m, H, dtype=rows.dtype, device=dev)
        pad[:rows.shape[0]] = rows
    flat = torch.empty(W * m, H, dtype=rows.dtype, device=dev)
    if timing is not None and rows.is_cuda:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    dist.all_gather_into_tensor(flat, pad, group=group)
    if timing is not None and rows.is_cuda:
        e1.record()
        timing.append((e0, e1, W * m * H * rows.element_size()))
    g.flat = flat
    return g


def all_gather_packed(rows: torch.Tensor, lens: torch.Tensor, group=None, timing=None, return_host: bool = False,
                      n_global: Optional[int] = None):
    """``gather_packed`` + ``select(None)``: ``(rows_global [sum over all utterances, H], lens_global [n_utts])`` in
    GLOBAL utterance order (utterance i lives on rank i % W)."""
    g = gather_packed(rows, lens, n_global=n_global, group=group, timing=timing)
    if g.W == 1:
        return (rows, lens, g.lens_host.tolist()) if return_host else (rows, lens)
    rows_g, lens_g = g.select(None)
    lens_g = lens_g.to(dtype=lens.dtype, device=lens.device)
    return (rows_g, lens_g, g.lens_host.tolist()) if return_host else (rows_g, lens_g)


def packed_inference_step(bridge, raw_encoder_out: torch.Tensor, raw_encoder_out_lens: torch.Tensor,
                          input_ids_global: torch.Tensor, attention_mask_global: torch.Tensor,
                          prompt_lens: Sequence[int], group=None, timing=None):
    """One inference step of BASELINE.json configs[3] on this rank: compress + project the rank's utterance shard
    (``TasuBridge.compress_project_async``), all-gather compressed lengths and packed rows (``gather_packed``), re-deal
    the global batch into length-homogeneous per-rank batches (``length_grouped_partition``), pick this rank's
    utterances straight out of the all-gather buffer (``PackedGather.select``) and splice them (``TasuBridge.splice``).

    ``input_ids_global`` / ``attention_mask_global`` ``[n_global, S]`` (left-padded prompts, identical on every rank,
    on the device), ``prompt_lens`` their token counts on the host.  Returns ``(splice outputs, info)`` with
    ``info = {sel, lens_host, valid_tokens, padded_tokens, rows_global}``.
    Host synchronisations: the plan header (overlapped with the tail kernels), the gathered lengths, the splice header."""
    import numpy as np
    rank, W = world()
    n_global = input_ids_global.shape[0]
    dev = raw_encoder_out.device
    pend = bridge.compress_project_async(raw_encoder_out, raw_encoder_out_lens)
    rows, lens, _ = pend.finish()
    g = gather_packed(rows, lens, n_global=n_global, group=group, timing=timing)
    plen = np.asarray(prompt_lens, dtype=np.int64)
    tot = plen + g.lens_host - 1                                   # spliced length of every utterance (one <speech> each)
    share = length_grouped_partition(tot.tolist(), W)
    sel = sorted(share[rank])
    info = {"sel": sel, "lens_host": g.lens_host, "rows_global": int(g.lens_host.sum()),
            "valid_tokens": int(tot[sel].sum()) if sel else 0, "padded_tokens": 0}
    if not sel:
        return None, info
    rows_sel, lens_sel = g.select(sel)
    sel_t = torch.as_tensor(sel, dtype=torch.int64).to(dev, non_blocking=True)
    cut = int(input_ids_global.shape[1] - int(plen[sel].max()))    # drop the columns that are padding for this group
    ids_s = input_ids_global.index_select(0, sel_t)[:, cut:].contiguous()
    mask_s = attention_mask_global.index_select(0, sel_t)[:, cut:].contiguous()
    out = bridge.splice(rows_sel, lens_sel, ids_s, mask_s)
    info["padded_tokens"] = int(out[0].shape[0] * out[0].shape[1])
    return out, info


# ---- gradient all-reduce overlapped with the backward (token-row projector) -------------------------------------------
_OVERLAP = {"on": False, "group": None, "pending": {}, "params": None}


def enable_overlapped_allreduce(on: bool = True, group=None, params: Optional[Sequence[torch.nn.Parameter]] = None):
    """Data-parallel training without DeepSpeed: when on, the token-row backward starts the all-reduce of its W1 half
    (dgamma | dbeta | dW1 | db1 = 94 % of the bytes) as soon as it is enqueued, so NCCL runs under the W2 half of the
    backward; ``allreduce_gradients`` then only sends the remainder and waits.  ``params`` (required when on): the
    parameters whose gradients that backward produces — the early all-reduce is only started when none of them already
    holds a gradient (with gradient accumulation autograd adds the new gradient INTO the old storage, and a half that
    was already summed over the ranks must not be reduced again).  Off (default): nothing is communicated inside
    backward (what DeepSpeed / an external reducer expects)."""
    if on and params is None:
        raise ValueError("enable_overlapped_allreduce(True) needs the parameters of the overlapped backward")
    _drain_pending()
    _OVERLAP["on"], _OVERLAP["group"] = bool(on), group
    _OVERLAP["params"] = list(params) if params is not None else None


def _drain_pending():
    """Wait for and forget every early all-reduce that no ``allreduce_gradients`` call picked up."""
    for work, _ in _OVERLAP["pending"].values():
        work.wait()
    _OVERLAP["pending"].clear()


def overlap_hook():
    """Callback for ``ops.tokrow_linear_silu_bwd(between=...)`` or None when overlap is off / single process / the
    parameters already hold gradients (accumulation: the whole gradient is reduced once, after the backward)."""
    rank, W = world()
    if not _OVERLAP["on"] or W == 1:
        return None
    _drain_pending()                                             # a backward whose reduction was never finished
    if any(p.grad is not None for p in _OVERLAP["params"]):
        return None

    def between(flat, n_first):
        work = dist.all_reduce(flat[:n_first], op=dist.ReduceOp.SUM, group=_OVERLAP["group"], async_op=True)
        _OVERLAP["pending"][flat.untyped_storage().data_ptr()] = (work, n_first)
    return between


def _shared_flat(grads):
    """The whole storage as one 1-D tensor if every gradient is a view of the same fp32 storage (the token-row
    backward hands out views of one flat buffer), else None."""
    if not grads:
        return None
    st = grads[0].untyped_storage()
    if any(g.dtype != torch.float32 or g.untyped_storage().data_ptr() != st.data_ptr() for g in grads):
        return None
    covered = sum(g.numel() for g in grads) * 4
    if covered < 0.9 * st.nbytes():                              # views of something much larger: not ours
        return None
    return torch.empty(0, dtype=torch.float32, device=grads[0].device).set_(st)


def allreduce_gradients(params: Sequence[torch.nn.Parameter], bucket_bytes: int = 64 << 20, average: bool = True,
                        group=None, async_op: bool = False, wire_dtype: Optional[torch.dtype] = None):
    """Gradient all-reduce of the projector parameters (54 512 062 params = 218 MB fp32 for linear-silu; replaces the
    ZeRO-2 reduce-scatter of conf/ds_config.json:15-21).  The token-row backward writes every gradient into ONE flat
    buffer, which is reduced in place as a single message; other gradients go in buckets sized for launch latency
    (NVSwitch gives every GPU full bandwidth to every peer, so not for link count).
    ``wire_dtype=torch.bfloat16`` (flat-buffer case, CUDA): the message is sent as bf16 — half the bytes; the sum over
    ranks is formed by NCCL in bf16, the result is widened and averaged back into the fp32 buffer.  Default (None): fp32
    on the wire, as the reference.  Returns the list of (work, flat, grads) handles when async."""
    rank, W = world()
    if W == 1:
        return []
    grads = [p.grad for p in params if p.grad is not None]
    flat = _shared_flat(grads)
    if flat is not None:                                         # one in-place message, no flatten / copy-back
        early = _OVERLAP["pending"].pop(flat.untyped_storage().data_ptr(), None)
        if early is None and wire_dtype == torch.bfloat16 and flat.is_cuda and not async_op:
            from . import ops
            wire = torch.empty(flat.numel(), dtype=torch.bfloat16, device=flat.device)
            ops.flat_scale_cast(flat, wire, 1.0)
            dist.all_reduce(wire, op=dist.ReduceOp.SUM, group=group)
            ops.flat_scale_cast(wire, flat, 1.0 / W if average else 1.0)
            return []
        if early is not None:                                    # the W1 half is already in flight (overlap_hook)
            work0, n_first = early
            work = dist.all_reduce(flat[n_first:], op=dist.ReduceOp.SUM, group=group, async_op=True)
            work0.wait()
        elif average and flat.is_cuda and not async_op:
            # NCCL averages inside the collective (ncclAvg): no second pass over the 218 MB buffer for the division
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
            return []
        else:
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True)
        handles = [(work, flat, [])]
        if async_op:
            return handles
        finish_allreduce(handles, W if average else 1)
        return []
    _drain_pending()                                             # gradients were accumulated: nothing early applies
    handles, bucket, size = [], [], 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True)
        handles.append((work, flat, bucket))
        bucket, size = [], 0

    for g in grads:
        bucket.append(g)
        size += g.numel() * g.element_size()
        if size >= bucket_bytes:
            flush()
    flush()
    if async_op:
        return handles
    finish_allreduce(handles, W if average else 1)
    return []


def reduce_scatter_gradients(params: Sequence[torch.nn.Parameter], average: bool = True, group=None):
    """ZeRO-2-style gradient exchange — what the reference's DeepSpeed configuration does with the projector gradients
    (conf/ds_config.json:15-21: ``"stage": 2, "reduce_scatter": true``): every rank receives the summed (averaged)
    gradient of ITS contiguous 1/W shard of the flat gradient buffer, half the bytes of an all-reduce on the wire; the
    sharded optimizer then updates that shard and all-gathers the parameters (DeepSpeed's side, out of scope here).
    Needs the gradients to be views of one flat fp32 buffer whose length is a multiple of W (the token-row backward's
    layout: 64-element aligned slices), else they are flattened (one copy) and zero padded.
    Returns ``(shard fp32 [n / W], (first, last) element range of the shard in the flat parameter order, flat length)``;
    single process: the whole buffer."""
    rank, W = world()
    grads = [p.grad for p in params if p.grad is not None]
    flat = _shared_flat(grads)
    if flat is None:
        flat = torch.cat([g.reshape(-1).float() for g in grads]) if grads else torch.zeros(0)
    n = flat.numel()
    if W == 1:
        return flat, (0, n), n
    _drain_pending()
    if n % W:
        flat = torch.cat([flat, flat.new_zeros(W - n % W)])
    per = flat.numel() // W
    if flat.is_cuda:
        shard = torch.empty(per, dtype=flat.dtype, device=flat.device)
        dist.reduce_scatter_tensor(shard, flat, op=dist.ReduceOp.AVG if average else dist.ReduceOp.SUM, group=group)
    else:                                                        # gloo (CPU tests) has no reduce-scatter
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        shard = flat[rank * per:(rank + 1) * per].clone()
        if average:
            shard /= W
    return shard, (rank * per, min((rank + 1) * per, n)), n


def finish_allreduce(handles, divisor: int):
    for work, flat, bucket in handles:
        work.wait()
        if divisor != 1:
            flat.div_(divisor)
        o = 0
        for g in bucket:
            n = g.numel()
            g.copy_(flat[o:o + n].view_as(g))
            o += n
