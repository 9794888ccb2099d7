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


def all_gather_packed(rows: torch.Tensor, lens: torch.Tensor, group=None, timing=None, return_host: bool = False):
    """Gather every rank's packed compressed rows ``[sum M_b, H]`` and lengths.

    Returns ``(rows_global [sum over all utterances, H], lens_global [n_utts])`` in GLOBAL utterance
    order (utterance i lives on rank i % W).  Two small collectives for the lengths, ONE host read of them,
    ONE padded-to-max all-gather of the payload (on NVSwitch a flat all-gather is bandwidth-optimal, no
    topology-aware ring needed) and one row-gather kernel that puts the rows in global order."""
    rank, W = world()
    if W == 1:
        return (rows, lens, lens.tolist()) if return_host else (rows, lens)
    all_lens = all_gather_lengths(lens, group)
    lens_host = [l.cpu() for l in all_lens]                      # the single device→host hand-off
    totals = [int(l.sum()) for l in lens_host]
    m = max(totals + [1])
    H = rows.shape[1]
    pad = torch.empty(m, H, dtype=rows.dtype, device=rows.device)
    pad[:rows.shape[0]] = rows
    flat = torch.empty(W * m, H, dtype=rows.dtype, device=rows.device)
    if timing is not None and rows.is_cuda:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    dist.all_gather_into_tensor(flat, pad, group=group)
    if timing is not None and rows.is_cuda:
        e1.record()
        timing.append((e0, e1, W * m * H * rows.element_size()))
    import numpy as np
    n = sum(l.numel() for l in lens_host)
    # global utterance i lives on rank i % W at local index i // W: its rows are flat[r*m + off_r[j] : ... + len]
    lens_np = [l.numpy().astype(np.int64) for l in lens_host]
    offs_np = [np.concatenate([[0], np.cumsum(l)]) for l in lens_np]
    gi = np.arange(n)
    r_of, j_of = gi % W, gi // W
    glens_np = np.zeros(n, dtype=np.int64)
    starts = np.zeros(n, dtype=np.int64)
    for r in range(W):
        sel = r_of == r
        glens_np[sel] = lens_np[r][j_of[sel]]
        starts[sel] = r * m + offs_np[r][j_of[sel]]
    idx = torch.from_numpy(concat_ranges(starts, glens_np))
    glens = glens_np.tolist()
    if rows.is_cuda:
        idx = idx.pin_memory().to(rows.device, non_blocking=True)
    rows_g = _gather_rows(flat, idx)
    lens_g = torch.tensor(glens, dtype=lens.dtype).to(lens.device)
    return (rows_g, lens_g, glens) if return_host else (rows_g, lens_g)


# ---- gradient all-reduce overlapped with the backward (token-row projector) -------------------------------------------
_OVERLAP = {"on": False, "group": None, "pending": {}}


def enable_overlapped_allreduce(on: bool = True, group=None):
    """Data-parallel training without DeepSpeed: when on, the token-row backward starts the all-reduce of its W1 half
    (dgamma | dbeta | dW1 | db1 = 94 % of the bytes) as soon as it is enqueued, so NCCL runs under the W2 half of the
    backward; ``allreduce_gradients`` then only sends the remainder and waits.  Off (default): nothing is communicated
    inside backward (what DeepSpeed / an external reducer expects)."""
    _OVERLAP["on"], _OVERLAP["group"] = bool(on), group


def overlap_hook():
    """Callback for ``ops.tokrow_linear_silu_bwd(between=...)`` or None when overlap is off / single process."""
    rank, W = world()
    if not _OVERLAP["on"] or W == 1:
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
                        group=None, async_op: bool = False):
    """Bucketed gradient all-reduce of the projector parameters (54 512 062 params = 218 MB fp32 for
    linear-silu).  Buckets are sized for launch latency/overlap, not link count: NVSwitch gives every
    GPU full bandwidth to every peer.  Returns the list of (work, flat, grads) handles when async."""
    rank, W = world()
    if W == 1:
        return []
    grads = [p.grad for p in params if p.grad is not None]
    flat = _shared_flat(grads)
    if flat is not None:                                         # one in-place message, no flatten / copy-back
        early = _OVERLAP["pending"].pop(flat.untyped_storage().data_ptr(), None)
        if early is not None:                                    # the W1 half is already in flight (overlap_hook)
            work0, n_first = early
            work = dist.all_reduce(flat[n_first:], op=dist.ReduceOp.SUM, group=group, async_op=True)
            work0.wait()
        else:
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True)
        handles = [(work, flat, [])]
        if async_op:
            return handles
        finish_allreduce(handles, W if average else 1)
        return []
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
