"""Thin torch-tensor wrappers over the C ABI (one function per entry point).

PyTorch is plumbing here: it owns device memory and the current stream; every byte of the
bridge's arithmetic happens inside libtasu_bridge.so.  All wrappers enqueue on
``torch.cuda.current_stream()`` and never synchronise.
"""
from typing import Optional

import torch

from . import _lib as L

V_ALIGN = 64          # leading dimension padding (elements) of buffers this package owns
COUNTERS = {"launches": 0}   # kernels of libtasu_bridge.so launched through this module (bench's gpu_launches)


def _count(n: int = 1):
    COUNTERS["launches"] += n


# torch.cuda.current_stream() builds a Stream object behind three layers of Python (4 us; ~20 calls per bridge call on the
# launch-bound host path): ask the C++ side for the raw handle of the current device's current stream instead
_RAW_STREAM = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_GET_DEVICE = getattr(torch._C, "_cuda_getDevice", None)


def _stream():
    if _RAW_STREAM is not None and _GET_DEVICE is not None:
        return _RAW_STREAM(_GET_DEVICE())
    return torch.cuda.current_stream().cuda_stream


def _rows(dev, dtype, n: int, k: int):
    """``k`` arrays of ``n`` 4-byte elements from ONE allocation when every row stays 16-byte aligned (the per-call host
    path is allocation- and launch-bound at small batches: one ``torch.empty`` + one ``unbind`` instead of ``k`` calls)."""
    if n % 4 == 0 and n > 0:
        return torch.empty((k, n), dtype=dtype, device=dev).unbind(0)
    return tuple(torch.empty(max(n, 1), dtype=dtype, device=dev) for _ in range(k))


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return L.F32
    if t.dtype == torch.bfloat16:
        return L.BF16
    raise TypeError("tasu bridge supports float32 and bfloat16 tensors, got %s" % t.dtype)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise L.TasuError("tasu bridge ops need CUDA tensors (no CPU fallback exists)")


def set_option(option: int, value: int) -> None:
    """tasu_set_option: run-time switches of the library (``_lib.OPT_*``); 0 = the validated default path."""
    L.check(L.lib().tasu_set_option(int(option), int(value)), "tasu_set_option")


def get_option(option: int) -> int:
    v = L.lib().tasu_get_option(int(option))
    if v == -1 and not 0 <= int(option) < L.OPT_COUNT:
        L.check(v, "tasu_get_option")
    return v


def pad_to(n: int, a: int = V_ALIGN) -> int:
    return (n + a - 1) // a * a


def view3(x: torch.Tensor):
    """(batch_stride, row_stride) of a [B,T,V] view whose last dim is contiguous."""
    if x.dim() != 3:
        raise ValueError("expected a [B, T, V] tensor")
    if x.shape[-1] > 1 and x.stride(-1) != 1:
        x = x.contiguous()
    return x, x.stride(0), x.stride(1)


class FrameStats:
    """``dec_max`` / ``dec_sum`` (exact-decision mode): the softmax normalisers the collapse plan uses instead of
    ``row_max`` / ``row_sumexp`` (fp32-recomputed for the refined frames)."""
    __slots__ = ("argmax", "x_blank", "row_max", "row_sumexp", "row_sumexp2", "gmax", "kind", "B", "T", "dec_max", "dec_sum")

    def __init__(self):
        self.dec_max = self.dec_sum = None


def frame_stats(x: torch.Tensor, input_kind: int, blank_id: int, lens: Optional[torch.Tensor] = None) -> FrameStats:
    """tasu_frame_stats on a [B,T,V] view (any batch/row stride, last dim contiguous)."""
    _need_cuda(x, lens)
    x, bs, rs = view3(x)
    B, T, V = x.shape
    dev = x.device
    st = FrameStats()
    st.kind, st.B, st.T = input_kind, B, T
    st.argmax = torch.empty(B * T, dtype=torch.int32, device=dev)
    st.x_blank = torch.empty(B * T, dtype=torch.float32, device=dev)
    st.row_max = torch.empty(B * T, dtype=torch.float32, device=dev)
    st.row_sumexp = torch.empty(B * T, dtype=torch.float32, device=dev) if input_kind == L.INPUT_LOGITS else None
    st.row_sumexp2 = None
    st.gmax = torch.empty(1, dtype=torch.int32, device=dev)
    if lens is not None:
        lens = lens.to(torch.int64)
    L.check(L.lib().tasu_frame_stats(x.data_ptr(), _dt(x), input_kind, B, T, V, bs, rs, blank_id, _ptr(lens),
                                     st.argmax.data_ptr(), st.x_blank.data_ptr(), st.row_max.data_ptr(),
                                     _ptr(st.row_sumexp), st.gmax.data_ptr(), _stream()), "tasu_frame_stats")
    _count(1)
    return st


def ctc_head_stats(x_bf16: torch.Tensor, w_bf16: torch.Tensor, bias: Optional[torch.Tensor], B: int, T: int,
                   n_prefix: int, V: int, K: int, blank_id: int) -> FrameStats:
    """Fused CTC head + softmax statistics (tasu_ctc_head_stats): x_bf16 [B*(T+P), ld], w_bf16 [V, ld]."""
    _need_cuda(x_bf16, w_bf16, bias)
    dev = x_bf16.device
    st = FrameStats()
    st.kind, st.B, st.T = L.INPUT_LOGITS, B, T
    st.x_blank, st.row_max, st.row_sumexp, st.row_sumexp2 = _rows(dev, torch.float32, B * T, 4)
    st.argmax = torch.empty(B * T, dtype=torch.int32, device=dev)
    st.gmax = torch.empty(1, dtype=torch.int32, device=dev)
    nbytes = L.lib().tasu_ctc_head_stats_workspace(B, T, n_prefix)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    L.check(L.lib().tasu_ctc_head_stats(x_bf16.data_ptr(), x_bf16.stride(0), w_bf16.data_ptr(), w_bf16.stride(0),
                                        _ptr(bias), B, T, n_prefix, V, K, blank_id, st.argmax.data_ptr(),
                                        st.x_blank.data_ptr(), st.row_max.data_ptr(), st.row_sumexp.data_ptr(),
                                        st.row_sumexp2.data_ptr(), ws.data_ptr(), nbytes, _stream()),
            "tasu_ctc_head_stats")
    _count(2)
    return st


_FP_ARGS = {}       # (address, bytes) of every tensor → the ctypes argument arrays (the call sits on the per-batch host path)


def fingerprint(tensors, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """tasu_fingerprint of up to 8 device tensors → int64[1] (``out`` may be a pinned host slot; no sync here)."""
    import ctypes
    ts = [t for t in tensors if t is not None]
    if len(ts) > 8:
        raise ValueError("fingerprint takes at most 8 tensors per call")
    if not all(t.is_cuda for t in ts):
        _need_cuda(*ts)
    if not all(t.is_contiguous() for t in ts):
        ts = [t.detach().contiguous() for t in ts]
    key = tuple((t.data_ptr(), t.numel() * t.element_size()) for t in ts)
    args = _FP_ARGS.get(key)
    if args is None:
        n = len(ts)
        args = ((ctypes.c_void_p * max(n, 1))(*[k[0] for k in key]), (ctypes.c_int64 * max(n, 1))(*[k[1] for k in key]), n)
        if len(_FP_ARGS) > 64:
            _FP_ARGS.clear()
        _FP_ARGS[key] = args
    if out is None:
        out = torch.empty(1, dtype=torch.int64, device=ts[0].device)
    L.check(L.lib().tasu_fingerprint(ctypes.addressof(args[0]), ctypes.addressof(args[1]), args[2], out.data_ptr(), _stream()),
            "tasu_fingerprint")
    _count(1)
    return out


def row_norm_max(w: torch.Tensor) -> torch.Tensor:
    """max_r ||w_r||_2 as a device uint32[1] in the order-preserving encoding (tasu_row_norm_max)."""
    _need_cuda(w)
    if w.stride(-1) != 1:
        w = w.contiguous()
    out = torch.empty(1, dtype=torch.int32, device=w.device)
    L.check(L.lib().tasu_row_norm_max(w.data_ptr(), _dt(w), w.shape[0], w.shape[1], w.stride(0), out.data_ptr(), _stream()),
            "tasu_row_norm_max")
    _count(1)
    return out


ERR_SCALE_F32_INPUT = 1.05 * 2.0 ** -8     # x and W both rounded to bf16 by the fused head (+5 % for the fp32 accumulation)
ERR_SCALE_BF16_INPUT = 1.05 * 2.0 ** -9    # x was GIVEN in bf16: only W is rounded


def cast_rows_sumsq(src: torch.Tensor, dst_stride: int):
    """fp32 [rows, cols] → (bf16 [rows, dst_stride] with zeroed pad columns, squared row norms fp32 [rows])."""
    _need_cuda(src)
    if src.dtype != torch.float32 or src.dim() != 2:
        raise TypeError("cast_rows_sumsq expects a 2-D float32 tensor")
    if src.shape[1] > 1 and src.stride(1) != 1:
        src = src.contiguous()
    rows, cols = src.shape
    dst = torch.empty(rows, dst_stride, dtype=torch.bfloat16, device=src.device)
    sumsq = torch.empty(max(rows, 1), dtype=torch.float32, device=src.device)
    L.check(L.lib().tasu_cast_rows_sumsq(src.data_ptr(), rows, cols, src.stride(0) if rows > 1 else cols, dst.data_ptr(),
                                         dst_stride, sumsq.data_ptr(), _stream()), "tasu_cast_rows_sumsq")
    _count(1)
    return dst, sumsq


def refine_ambiguous_frames(st: FrameStats, lens: torch.Tensor, x_rows: torch.Tensor, w_f32: torch.Tensor,
                            bias: Optional[torch.Tensor], w_norm_max: torch.Tensor, T: int, n_prefix: int, V: int,
                            blank_id: int, threshold: float, x_sumsq: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Exact-decision mode (csrc/refine.cu): list the frames whose greedy decisions (argmax, ps-slm.py:265; strict fp32
    blank threshold, :295-297) lie inside the rounding error bound of the bf16 head, recompute exactly those with fp32
    FMAs from the fp32 weights, and publish the decision statistics in ``st`` (``argmax`` / ``x_blank`` in place,
    ``dec_max`` / ``dec_sum`` new) for ``collapse_plan``.  ``x_rows`` = [B*(T+P), K] encoder rows as GIVEN (fp32 or bf16);
    ``x_sumsq``: their squared norms when already known (``cast_rows_sumsq``).
    The list holds one slot per frame: nothing is capped or dropped.  Returns the device counter (int32[1]); no sync."""
    _need_cuda(x_rows, w_f32, bias)
    dev = st.argmax.device
    n = st.B * T
    if x_rows.stride(-1) != 1:
        x_rows = x_rows.contiguous()
    K = x_rows.shape[1]
    lists = torch.empty(2, max(n, 1), dtype=torch.int32, device=dev)
    count = torch.empty(1, dtype=torch.int32, device=dev)
    dec = torch.empty(2, max(n, 1), dtype=torch.float32, device=dev)
    lib = L.lib()
    lens = lens.to(device=dev, dtype=torch.int64)
    err = ERR_SCALE_BF16_INPUT if x_rows.dtype == torch.bfloat16 else ERR_SCALE_F32_INPUT
    L.check(lib.tasu_flag_ambiguous_frames(st.argmax.data_ptr(), st.x_blank.data_ptr(), st.row_max.data_ptr(),
                                           st.row_sumexp.data_ptr(), _ptr(st.row_sumexp2), lens.data_ptr(),
                                           x_rows.data_ptr(), _dt(x_rows), x_rows.stride(0), K, _ptr(x_sumsq),
                                           w_norm_max.data_ptr(),
                                           float(err), st.B, T, n_prefix, blank_id, float(threshold), dec[0].data_ptr(),
                                           dec[1].data_ptr(), lists[0].data_ptr(), lists[1].data_ptr(), count.data_ptr(),
                                           _stream()), "tasu_flag_ambiguous_frames")
    nbytes = lib.tasu_ctc_head_refine_workspace(n)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    L.check(lib.tasu_ctc_head_refine(x_rows.data_ptr(), _dt(x_rows), x_rows.stride(0), w_f32.data_ptr(), w_f32.stride(0),
                                     _ptr(bias), V, K, blank_id, lists[0].data_ptr(), lists[1].data_ptr(), count.data_ptr(),
                                     n, st.argmax.data_ptr(), st.x_blank.data_ptr(), dec[0].data_ptr(), dec[1].data_ptr(),
                                     ws.data_ptr(), nbytes, _stream()), "tasu_ctc_head_refine")
    st.dec_max, st.dec_sum = dec[0], dec[1]
    _count(3)
    return count


def gather_kept_rows(x_bf16: torch.Tensor, B: int, T: int, n_prefix: int, K: int, V: int, plan: "CollapsePlan",
                     st: FrameStats, n_frames: int, n_out: int, ln_eps: float = 1e-5):
    """→ (xg bf16 [n_frames, pad64(K)], g_max, g_inv_sum [n_frames], pk_len, tail_src int32 [n_out],
    ln_mean, ln_rstd [n_out]).  Rows [0, n_out) are the first frames of the packed candidates."""
    dev = x_bf16.device
    ldg = pad_to(K)
    cap = (max(n_frames, 1) + 2047) // 2048 * 2048          # few distinct sizes → allocator cache hits
    xg = torch.empty(cap, ldg, dtype=torch.bfloat16, device=dev)[:max(n_frames, 1)]
    g_max, g_inv = _rows(dev, torch.float32, n_frames, 2)
    pk_len, tail_src = _rows(dev, torch.int32, n_out, 2)
    multi = torch.empty(max(n_out, 1) + 1, dtype=torch.int32, device=dev)       # [0] = count, [1:] = row list
    mean, rstd = _rows(dev, torch.float32, n_out, 2)
    L.check(L.lib().tasu_gather_kept_rows(x_bf16.data_ptr(), x_bf16.stride(0), B, T, n_prefix, K, V,
                                          plan.seg_start.data_ptr(), plan.seg_len.data_ptr(), plan.seg_foff.data_ptr(),
                                          plan.row_off.data_ptr(), plan.frame_off.data_ptr(), st.row_max.data_ptr(),
                                          st.row_sumexp.data_ptr(), _ptr(st.row_sumexp2), n_frames, n_out, xg.data_ptr(), ldg,
                                          g_max.data_ptr(), g_inv.data_ptr(), pk_len.data_ptr(), tail_src.data_ptr(),
                                          multi[1:].data_ptr(), multi.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                          float(ln_eps), _stream()), "tasu_gather_kept_rows")
    _count(1)
    return xg, g_max, g_inv, pk_len, tail_src, multi, mean, rstd


def kept_frame_index(plan: "CollapsePlan", n_prefix: int, n_frames: int, n_out: int):
    """(frame_row int32 [n_frames] raw encoder row of every kept frame in natural order, seg_src int32 [n_out] first
    compact row of every packed candidate) — tasu_kept_frame_index."""
    dev = plan.seg_start.device
    frame_row = torch.empty(max(n_frames, 1), dtype=torch.int32, device=dev)[:n_frames]
    seg_src = torch.empty(max(n_out, 1), dtype=torch.int32, device=dev)[:n_out]
    L.check(L.lib().tasu_kept_frame_index(plan.seg_start.data_ptr(), plan.seg_len.data_ptr(), plan.seg_foff.data_ptr(),
                                          plan.row_off.data_ptr(), plan.frame_off.data_ptr(), plan.B, plan.T, n_prefix,
                                          n_frames, n_out, frame_row.data_ptr(), seg_src.data_ptr(), _stream()),
            "tasu_kept_frame_index")
    _count(1)
    return frame_row, seg_src


def pool_tail(probs: torch.Tensor, D: int, n_out: int, pk_len: torch.Tensor, tail_src: torch.Tensor,
              multi: Optional[torch.Tensor], ln_mean: torch.Tensor, ln_rstd: torch.Tensor, ln_eps: float = 1e-5):
    """``probs``: the compact [rows, ld] matrix; its row count is the capacity no access may exceed."""
    L.check(L.lib().tasu_pool_tail(probs.data_ptr(), probs.stride(0), D, n_out, probs.shape[0], pk_len.data_ptr(), tail_src.data_ptr(),
                                   multi[1:].data_ptr() if multi is not None else None, _ptr(multi),
                                   ln_mean.data_ptr(), ln_rstd.data_ptr(), float(ln_eps), _stream()), "tasu_pool_tail")
    _count(1)


class GroupedRows:
    """Kept frames in the grouped layout (include/tasu_bridge.h, step 2b'): ``xg`` = A operand of the kept-frame GEMM,
    ``g_max`` / ``g_inv`` its per-row softmax scalars, ``perm`` [cap_o] packed candidate → pooled row, ``lay`` the layout
    words on the device, ``mean`` / ``rstd`` / ``pk_len`` / ``tail_src`` by pooled row, ``multi`` pool_tail's work list."""
    __slots__ = ("xg", "g_max", "g_inv", "perm", "lay", "pk_len", "tail_src", "multi", "mean", "rstd", "cap_a", "cap_p")


def grouped_capacities(n_frames: int, n_out: int):
    """(A rows, pooled rows) that hold ANY batch of at most ``n_frames`` kept frames: region padding (3 x 127 rows) plus
    one zero row per 3-frame run (at most n_frames / 3) on the A side; the pooled matrix carries no zero rows."""
    cap_a = n_frames + n_frames // 3 + 512
    return cap_a, n_frames + 512        # pooled rows: N_out + at most 221 hole rows + the extra frames of the long runs


def gather_kept_rows_grouped(x_bf16: torch.Tensor, B: int, T: int, n_prefix: int, K: int, V: int, plan: "CollapsePlan",
                             st: FrameStats, n_frames: int, n_out: int, max_proj: int, ln_eps: float = 1e-5) -> GroupedRows:
    """tasu_group_plan + tasu_gather_kept_rows_grouped (two launches, no host synchronisation).  ``n_frames`` / ``n_out``
    are capacities (kept frames / packed candidates), ``max_proj`` the pooled rows the projector's buffers hold."""
    dev = x_bf16.device
    ldg = pad_to(K)
    g = GroupedRows()
    cap_a, cap_p = grouped_capacities(max(n_frames, 1), max(n_out, 1))
    cap_a = (cap_a + 2047) // 2048 * 2048                    # few distinct sizes → allocator cache hits
    g.cap_a, g.cap_p = cap_a, (cap_p + 2047) // 2048 * 2048
    g.xg = torch.empty(cap_a, ldg, dtype=torch.bfloat16, device=dev)
    g.g_max, g.g_inv, g.mean, g.rstd = _rows(dev, torch.float32, cap_a, 4)
    g.pk_len, g.tail_src = _rows(dev, torch.int32, cap_a, 2)
    ints = torch.empty(2 * B * T + 8 * B + L.GL_WORDS + max(n_out, 1) * 2 + 1, dtype=torch.int32, device=dev)
    slot, xoff = ints[:B * T], ints[B * T:2 * B * T]
    o = 2 * B * T
    cnt, base = ints[o:o + 4 * B], ints[o + 4 * B:o + 8 * B]
    o += 8 * B
    g.lay = ints[o:o + L.GL_WORDS]
    o += L.GL_WORDS
    g.perm = ints[o:o + max(n_out, 1)]
    g.multi = ints[o + max(n_out, 1):]                       # [0] = count, [1:] = pooled-row list
    L.check(L.lib().tasu_group_plan(plan.seg_len.data_ptr(), plan.new_lens.data_ptr(), B, T, slot.data_ptr(), xoff.data_ptr(),
                                    cnt.data_ptr(), base.data_ptr(), g.lay.data_ptr(), _ticket(dev)[2:3].data_ptr(), _stream()),
            "tasu_group_plan")
    L.check(L.lib().tasu_gather_kept_rows_grouped(
        x_bf16.data_ptr(), x_bf16.stride(0), B, T, n_prefix, K, V, plan.seg_start.data_ptr(), plan.seg_len.data_ptr(),
        plan.row_off.data_ptr(), slot.data_ptr(), xoff.data_ptr(), base.data_ptr(), g.lay.data_ptr(), st.row_max.data_ptr(),
        st.row_sumexp.data_ptr(), _ptr(st.row_sumexp2), cap_a, g.cap_p, max(n_out, 1), max_proj, g.xg.data_ptr(), ldg,
        g.g_max.data_ptr(), g.g_inv.data_ptr(), g.perm.data_ptr(), g.pk_len.data_ptr(), g.tail_src.data_ptr(),
        g.multi[1:].data_ptr(), g.multi.data_ptr(), g.mean.data_ptr(), g.rstd.data_ptr(), float(ln_eps), _stream()),
        "tasu_gather_kept_rows_grouped")
    _count(2)
    return g


def gemm_softmax_grouped(g: GroupedRows, w_bf16: torch.Tensor, bias: torch.Tensor, V: int, K: int, pooled: torch.Tensor,
                         ln_eps: float = 1e-5):
    """tasu_gemm_softmax_grouped + tasu_group_ln_finish: pooled probabilities of every kept candidate into ``pooled``
    [cap_p, pad64(V)] bf16 (class order, rows of long runs still per frame: ``pool_tail`` follows) and the LayerNorm
    statistics of the rows pooled in the epilogue."""
    _need_cuda(g.xg, w_bf16, pooled)
    n_parts = L.lib().tasu_gemm_softmax_grouped_parts(V)
    ldq = max(int(g.perm.numel()), 64) + 96                  # pooled rows of the G regions: at most n_out + 63 + 31
    q_part = torch.empty(n_parts, ldq, dtype=torch.float32, device=pooled.device)
    L.check(L.lib().tasu_gemm_softmax_grouped(g.xg.data_ptr(), g.xg.stride(0), w_bf16.data_ptr(), w_bf16.stride(0),
                                              pooled.data_ptr(), pooled.stride(0), g.cap_a, pooled.shape[0], V, K,
                                              bias.data_ptr(), g.g_inv.data_ptr(), g.g_max.data_ptr(), g.lay.data_ptr(),
                                              q_part.data_ptr(), ldq, _stream()), "tasu_gemm_softmax_grouped")
    L.check(L.lib().tasu_group_ln_finish(q_part.data_ptr(), ldq, n_parts, g.lay.data_ptr(), V, g.cap_p, g.mean.data_ptr(),
                                         g.rstd.data_ptr(), float(ln_eps), _stream()), "tasu_group_ln_finish")
    _count(2)
    return pooled


class CollapsePlan:
    __slots__ = ("seg_start", "seg_len", "seg_score", "seg_foff", "new_lens", "kept_frames", "row_off", "frame_off",
                 "header", "counts", "B", "T")


_TICKETS = {}


def _ticket(device) -> torch.Tensor:
    """Per (device, stream) int32[1] ticket of the fused planner kernels: zero-initialised once, handed back zeroed by
    every launch (stream-ordered reuse)."""
    key = (torch.device(device).index, torch.cuda.current_stream().cuda_stream)
    t = _TICKETS.get(key)
    if t is None:
        t = torch.zeros(4, dtype=torch.int32, device=device)
        _TICKETS[key] = t
    return t


def collapse_plan(st: FrameStats, lens: torch.Tensor, blank_id: int, threshold: float,
                  header: Optional[torch.Tensor] = None, want_scores: bool = False) -> CollapsePlan:
    """tasu_collapse_plan_scan (plan + scans + header in one launch). ``header`` (int64[>=4]) may be caller-provided
    so several plans share one device→host read."""
    B, T = st.B, st.T
    dev = st.argmax.device
    lens = lens.to(device=dev, dtype=torch.int64).contiguous()
    p = CollapsePlan()
    p.B, p.T = B, T
    p.seg_start, p.seg_len, p.seg_foff = _rows(dev, torch.int32, B * T, 3)
    p.seg_score = torch.empty(max(B * T, 1), dtype=torch.float32, device=dev) if want_scores else None
    p.new_lens = torch.empty(B, dtype=torch.int64, device=dev)
    p.kept_frames = torch.empty(max(B, 1), dtype=torch.int32, device=dev)
    p.frame_off = torch.empty(B + 1, dtype=torch.int32, device=dev)
    p.counts = torch.empty(4, dtype=torch.int32, device=dev)       # {N_out, max_len, kept_frames, 0} for device-side M
    p.row_off = torch.empty(B + 1, dtype=torch.int32, device=dev)
    # ``header`` may be pinned host memory: under UVA its pointer is valid on the device
    p.header = header if header is not None else torch.empty(L.CH_WORDS, dtype=torch.int64, device=dev)
    d_max = st.dec_max if st.dec_max is not None else st.row_max
    d_sum = st.dec_sum if st.dec_sum is not None else st.row_sumexp
    L.check(L.lib().tasu_collapse_plan_scan(st.argmax.data_ptr(), st.x_blank.data_ptr(), d_max.data_ptr(), _ptr(d_sum),
                                            st.gmax.data_ptr(), st.kind, lens.data_ptr(), B, T, blank_id, float(threshold),
                                            p.seg_start.data_ptr(), p.seg_len.data_ptr(), _ptr(p.seg_score),
                                            p.new_lens.data_ptr(), p.kept_frames.data_ptr(), p.seg_foff.data_ptr(),
                                            p.row_off.data_ptr(), p.frame_off.data_ptr(), p.header.data_ptr(),
                                            p.counts.data_ptr(), _ticket(dev)[0:1].data_ptr(), _stream()),
            "tasu_collapse_plan_scan")
    _count(1)
    return p


def segment_meanpool(feats: torch.Tensor, plan: CollapsePlan, layout: int, max_len: int, max_rows: int,
                     out: torch.Tensor, out_row_stride: int, softmax: Optional[FrameStats] = None,
                     ln_mean: Optional[torch.Tensor] = None, ln_rstd: Optional[torch.Tensor] = None,
                     ln_eps: float = 1e-5, seg_src: Optional[torch.Tensor] = None, feat_dim: Optional[int] = None):
    """``seg_src`` given → ``feats`` is the compact [F_kept, pitch] matrix of gather_kept_rows (2-D)."""
    _need_cuda(feats, out)
    if seg_src is not None:
        B, T, D = plan.B, plan.T, feat_dim
        bs, rs = 0, feats.stride(0)
    else:
        feats, bs, rs = view3(feats)
        B, T, D = feats.shape
    L.check(L.lib().tasu_segment_meanpool(
        feats.data_ptr(), _dt(feats), B, T, D, bs, rs,
        _ptr(softmax.row_max) if softmax is not None else None,
        _ptr(softmax.row_sumexp) if softmax is not None else None,
        plan.seg_start.data_ptr(), plan.seg_len.data_ptr(), plan.row_off.data_ptr(), _ptr(seg_src),
        layout, max_len, max_rows, out.data_ptr(), _dt(out), out_row_stride,
        _ptr(ln_mean), _ptr(ln_rstd), float(ln_eps), _stream()), "tasu_segment_meanpool")
    _count(1)
    return out


def cast_rows(src: torch.Tensor, dst_dtype: torch.dtype, dst_stride: Optional[int] = None, want_ln: bool = False,
              ln_eps: float = 1e-5):
    """[rows, cols] → dst dtype with pitch ``dst_stride`` (+ LayerNorm stats per row)."""
    _need_cuda(src)
    if src.dim() != 2:
        raise ValueError("cast_rows expects a 2-D tensor")
    if src.shape[1] > 1 and src.stride(1) != 1:
        src = src.contiguous()
    rows, cols = src.shape
    dst_stride = cols if dst_stride is None else dst_stride
    dst = torch.empty(rows, dst_stride, dtype=dst_dtype, device=src.device)
    mean = torch.empty(rows, dtype=torch.float32, device=src.device) if want_ln else None
    rstd = torch.empty(rows, dtype=torch.float32, device=src.device) if want_ln else None
    L.check(L.lib().tasu_cast_rows(src.data_ptr(), _dt(src), rows, cols, src.stride(0) if rows > 1 else cols,
                                   dst.data_ptr(), _dt(dst), dst_stride, _ptr(mean), _ptr(rstd), float(ln_eps),
                                   _stream()), "tasu_cast_rows")
    _count(1)
    return dst, mean, rstd


def fold_layernorm(w1: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, b1: Optional[torch.Tensor]):
    """(W1g bf16 [N, pad64(K)], colsum [N], dbias [N]) — LayerNorm folded into the first Linear."""
    _need_cuda(w1, gamma, beta, b1)
    N, K = w1.shape
    w1 = w1.float().contiguous()
    ld = pad_to(K)
    w1g = torch.empty(N, ld, dtype=torch.bfloat16, device=w1.device)
    colsum = torch.empty(N, dtype=torch.float32, device=w1.device)
    dbias = torch.empty(N, dtype=torch.float32, device=w1.device)
    L.check(L.lib().tasu_fold_layernorm(w1.data_ptr(), K, gamma.float().contiguous().data_ptr(),
                                        beta.float().contiguous().data_ptr(),
                                        _ptr(b1.float().contiguous()) if b1 is not None else None, N, K,
                                        w1g.data_ptr(), ld, colsum.data_ptr(), dbias.data_ptr(), _stream()),
            "tasu_fold_layernorm")
    _count(1)
    return w1g, colsum, dbias


def gemm_bf16_tn(A: torch.Tensor, Bw: torch.Tensor, M: int, N: int, K: int, out: torch.Tensor,
                 epilogue: int = L.EPI_NONE, bias: Optional[torch.Tensor] = None,
                 row_rstd: Optional[torch.Tensor] = None, row_mean: Optional[torch.Tensor] = None,
                 colsum: Optional[torch.Tensor] = None, simt: bool = False, m_dev: Optional[torch.Tensor] = None):
    """out[M,N] = epilogue(A[M,K] · Bw[N,K]^T); A/Bw bf16 with pitch = stride(0); out bf16|fp32.
    ``m_dev`` (int32 device scalar): live row count, M is then the allocated capacity."""
    _need_cuda(A, Bw, out)
    if A.dtype != torch.bfloat16 or Bw.dtype != torch.bfloat16:
        raise TypeError("GEMM operands must be bfloat16")
    fn = L.lib().tasu_gemm_bf16_tn_simt if simt else L.lib().tasu_gemm_bf16_tn
    lda = A.stride(0) if A.dim() == 2 and A.shape[0] > 1 else max(K, A.shape[-1])
    ldb = Bw.stride(0) if Bw.shape[0] > 1 else max(K, Bw.shape[-1])
    ldc = out.stride(0) if out.shape[0] > 1 else max(N, out.shape[-1])
    args = (A.data_ptr(), lda, Bw.data_ptr(), ldb, out.data_ptr(), _dt(out), ldc, M, N, K, epilogue,
            _ptr(bias), _ptr(row_rstd), _ptr(row_mean), _ptr(colsum))
    if simt:
        L.check(fn(*args, _stream()), "tasu_gemm_bf16_tn_simt")
    else:
        L.check(fn(*args, _ptr(m_dev), _stream()), "tasu_gemm_bf16_tn")
    _count(1)
    return out


_STREAMK_WS = {}


def streamk_workspace(device) -> torch.Tensor:
    """Workspace of the stream-K GEMM (flags + one fp32 partial tile per SM), zero-filled once and handed back with its
    flags zeroed by every launch.  One per (device, stream): launches on one stream are ordered, launches on different
    streams must not share flags."""
    idx = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    key = (idx, torch.cuda.current_stream(idx).cuda_stream)
    if key not in _STREAMK_WS:
        _STREAMK_WS[key] = torch.zeros(int(L.lib().tasu_gemm_streamk_workspace()), dtype=torch.uint8,
                                       device=torch.device("cuda", idx))
    return _STREAMK_WS[key]


def gemm_bf16_tn_streamk(A: torch.Tensor, Bw: torch.Tensor, M: int, N: int, K: int, out: torch.Tensor,
                         epilogue: int = L.EPI_NONE, bias: Optional[torch.Tensor] = None,
                         row_rstd: Optional[torch.Tensor] = None, row_mean: Optional[torch.Tensor] = None,
                         colsum: Optional[torch.Tensor] = None, m_dev: Optional[torch.Tensor] = None):
    """``gemm_bf16_tn`` for deep-K problems with the ragged last wave of tiles cut along K (tasu_gemm_bf16_tn_streamk;
    CTA pairs when TASU_OPT_GEMM_PAIR is on).  Deterministic; the cut tiles differ from the plain kernel in the last fp32 bits."""
    _need_cuda(A, Bw, out)
    if A.dtype != torch.bfloat16 or Bw.dtype != torch.bfloat16:
        raise TypeError("GEMM operands must be bfloat16")
    ws = streamk_workspace(A.device)
    lda = A.stride(0) if A.dim() == 2 and A.shape[0] > 1 else max(K, A.shape[-1])
    ldb = Bw.stride(0) if Bw.shape[0] > 1 else max(K, Bw.shape[-1])
    ldc = out.stride(0) if out.shape[0] > 1 else max(N, out.shape[-1])
    L.check(L.lib().tasu_gemm_bf16_tn_streamk(A.data_ptr(), lda, Bw.data_ptr(), ldb, out.data_ptr(), _dt(out), ldc, M, N, K,
                                              epilogue, _ptr(bias), _ptr(row_rstd), _ptr(row_mean), _ptr(colsum),
                                              _ptr(m_dev), ws.data_ptr(), ws.numel(), _stream()),
            "tasu_gemm_bf16_tn_streamk")
    _count(1)
    return out


def attn_softmax_pv(Q: torch.Tensor, table: torch.Tensor, N: int, V2: int, heads: int, dp: int, Z: torch.Tensor,
                    row_max: Optional[torch.Tensor] = None, row_inv: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``Z[:, h*dp:(h+1)*dp] = softmax(Q_h·K_hᵀ)·K_h`` for all heads in one launch, probabilities kept on the SM
    (tasu_attn_softmax_pv).  ``Q`` [N, heads*dp] bf16 (pre-scaled), ``table`` [V2, heads*dp] bf16, ``Z`` [N, heads*dp] fp32.
    ``row_max`` / ``row_inv`` [heads, N] fp32: statistics of the scores per head from a separate pass; without them the
    kernel finds the row maxima in a first sweep of its own (the default: no statistics pass at all)."""
    _need_cuda(Q, table, row_max, row_inv, Z)
    if Q.dtype != torch.bfloat16 or table.dtype != torch.bfloat16 or Z.dtype != torch.float32:
        raise TypeError("attn_softmax_pv: bf16 operands, fp32 output")
    ws, ws_bytes, launches = None, 0, 1
    if row_max is None and ATTN_KEY_SPLIT:
        # self-contained mode: key split so that the last wave of (row tile, head) items is full (tasu_attn_split_plan)
        import ctypes
        n_splits = ctypes.c_int(1)
        ws_bytes = int(L.lib().tasu_attn_split_plan(N, V2, heads, dp, ctypes.byref(n_splits)))
        if ws_bytes > 0:
            ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=Q.device)
            launches = 2
    L.check(L.lib().tasu_attn_softmax_pv_ws(Q.data_ptr(), Q.stride(0), table.data_ptr(), table.stride(0), N, V2, heads, dp,
                                            _ptr(row_max), _ptr(row_inv), row_max.stride(0) if row_max is not None else 0,
                                            Z.data_ptr(), Z.stride(0), _ptr(ws), ws_bytes, _stream()), "tasu_attn_softmax_pv_ws")
    _count(launches)
    return Z


ATTN_KEY_SPLIT = True       # False: one item per (row tile, head) — tasu_attn_softmax_pv as it was


ATTN_FUSED_WIDTHS = (64, 128, 192, 256)


def streamk_schedule(num_tiles: int, k_blocks: int, grid: int):
    """HOST: the stream-K schedule the kernel runs — ``(dp_tiles, [pieces of CTA 0, pieces of CTA 1, ...])`` with pieces
    ``(tile, kb0, kb1, kind, n_contrib)`` in processing order (tasu_gemm_streamk_schedule_host)."""
    import ctypes
    buf = (ctypes.c_int32 * 10)()
    dp = ctypes.c_int32(0)
    out = []
    for cta in range(grid):
        n = L.lib().tasu_gemm_streamk_schedule_host(num_tiles, k_blocks, grid, cta, ctypes.addressof(buf), ctypes.addressof(dp))
        if n < 0:
            L.check(n, "tasu_gemm_streamk_schedule_host")
        out.append([tuple(buf[5 * i:5 * i + 5]) for i in range(n)])
    return dp.value, out


def gemm_bf16_f32(A: torch.Tensor, a_mn_major: bool, Bw: torch.Tensor, b_mn_major: bool, M: int, N: int, K: int,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[M,N] fp32 = A · B with either operand K-major ([M|N, K]) or MN-major ([K, M|N]) in memory (tasu_gemm_bf16_f32)."""
    _need_cuda(A, Bw, out)
    if A.dtype != torch.bfloat16 or Bw.dtype != torch.bfloat16:
        raise TypeError("GEMM operands must be bfloat16")
    if out is None:
        out = torch.empty(M, pad_to(N, 4), dtype=torch.float32, device=A.device)[:, :N]
    L.check(L.lib().tasu_gemm_bf16_f32(A.data_ptr(), A.stride(0), int(a_mn_major), Bw.data_ptr(), Bw.stride(0), int(b_mn_major),
                                       out.data_ptr(), out.stride(0), M, N, K, _stream()), "tasu_gemm_bf16_f32")
    _count(1)
    return out


def softmax_rows(x2: torch.Tensor, V: int, st: FrameStats) -> torch.Tensor:
    """bf16 [rows, pad64(V)] = softmax(x2[:, :V]) with the row max / sum-exp of ``st`` (tasu_softmax_rows)."""
    _need_cuda(x2)
    rows = x2.shape[0]
    ld = pad_to(V)
    out = torch.empty(max(rows, 1), ld, dtype=torch.bfloat16, device=x2.device)[:rows]
    L.check(L.lib().tasu_softmax_rows(x2.data_ptr(), _dt(x2), x2.stride(0) if rows > 1 else max(V, x2.shape[-1]), rows, V,
                                      st.row_max.data_ptr(), st.row_sumexp.data_ptr(), out.data_ptr(), ld, _stream()),
            "tasu_softmax_rows")
    _count(1)
    return out


def gather_rows(src: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """out[i] = src[idx[i]] (idx int32; -1 → zero row) — tasu_gather_rows."""
    _need_cuda(src, idx)
    H = src.shape[1]
    n = idx.numel()
    out = torch.empty(max(n, 1), H, dtype=src.dtype, device=src.device)[:n]
    L.check(L.lib().tasu_gather_rows(src.data_ptr(), _dt(src), src.stride(0), idx.data_ptr(), n, H, out.data_ptr(), H, _stream()),
            "tasu_gather_rows")
    _count(1)
    return out


def sim_posterior_rows(tok: torch.Tensor, hot: torch.Tensor, base: torch.Tensor, V: int, out: torch.Tensor,
                       out_row_stride: int, dst_row: Optional[torch.Tensor] = None,
                       ln_mean: Optional[torch.Tensor] = None, ln_rstd: Optional[torch.Tensor] = None,
                       ln_eps: float = 1e-5):
    _need_cuda(tok, hot, base, out)
    L.check(L.lib().tasu_sim_posterior_rows(tok.data_ptr(), hot.data_ptr(), base.data_ptr(), _ptr(dst_row),
                                            tok.numel(), V, out.data_ptr(), _dt(out), out_row_stride,
                                            _ptr(ln_mean), _ptr(ln_rstd), float(ln_eps), _stream()),
            "tasu_sim_posterior_rows")
    _count(1)
    return out


class SplicePlan:
    __slots__ = ("rowstat", "new_pos", "text_prefix", "slot_ord", "slot_base", "audio_off", "header", "left_padding",
                 "B", "S", "n_audio", "mask_dtype", "speech_id", "input_ids", "attention_mask", "audio_dest", "row_src")


def _mask_arg(attention_mask: torch.Tensor):
    if attention_mask.dtype == torch.bool:
        return attention_mask.contiguous(), 0
    if attention_mask.dtype == torch.uint8:
        return attention_mask.contiguous(), 0
    return attention_mask.to(torch.int64).contiguous(), 1


def splice_rowstat(input_ids: torch.Tensor, attention_mask: torch.Tensor, speech_id: int) -> SplicePlan:
    _need_cuda(input_ids, attention_mask)
    B, S = input_ids.shape
    dev = input_ids.device
    p = SplicePlan()
    p.B, p.S, p.speech_id = B, S, int(speech_id)
    p.input_ids = input_ids.to(torch.int64).contiguous()
    p.attention_mask, p.mask_dtype = _mask_arg(attention_mask)
    p.rowstat = torch.empty(max(B, 1), 8, dtype=torch.int32, device=dev)
    L.check(L.lib().tasu_splice_rowstat(p.input_ids.data_ptr(), p.attention_mask.data_ptr(), p.mask_dtype, B, S,
                                        p.speech_id, p.rowstat.data_ptr(), _stream()), "tasu_splice_rowstat")
    _count(1)
    return p


def splice_plan(p: SplicePlan, num_audio: torch.Tensor, div_k: int = 1, header: Optional[torch.Tensor] = None):
    dev = p.input_ids.device
    num_audio = num_audio.to(device=dev, dtype=torch.int64).contiguous()
    B, S = p.B, p.S
    p.n_audio = num_audio.numel()
    p.new_pos = torch.empty(max(B * S, 1), dtype=torch.int32, device=dev)
    p.text_prefix = torch.empty(max(B * S, 1), dtype=torch.int32, device=dev)
    p.slot_ord = torch.empty(max(B * S, 1), dtype=torch.int32, device=dev)
    p.slot_base = torch.empty(max(B, 1), dtype=torch.int32, device=dev)
    p.audio_off = torch.empty(p.n_audio + 1, dtype=torch.int32, device=dev)
    p.header = header if header is not None else torch.empty(L.SH_WORDS, dtype=torch.int64, device=dev)
    L.check(L.lib().tasu_splice_plan_header(p.input_ids.data_ptr(), p.attention_mask.data_ptr(), p.mask_dtype, B, S,
                                            p.speech_id, num_audio.data_ptr(), p.n_audio, div_k, p.rowstat.data_ptr(),
                                            p.new_pos.data_ptr(), p.text_prefix.data_ptr(), p.slot_ord.data_ptr(),
                                            p.header.data_ptr(), p.slot_base.data_ptr(), p.audio_off.data_ptr(),
                                            _ticket(dev)[1:2].data_ptr(), _stream()), "tasu_splice_plan_header")
    _count(1)
    return p


def splice_scatter(p: SplicePlan, spliced_len: int, text_src: torch.Tensor, text_mode: int,
                   audio_rows: torch.Tensor, audio_layout: int, audio_max_len: int,
                   labels: Optional[torch.Tensor], pad_id: int, ignore_id: int, want_ids: bool = True,
                   left_padding: Optional[int] = None, want_audio_dest: bool = False,
                   audio_perm: Optional[torch.Tensor] = None):
    """Row map + one copy pass → (emb [B,S',H], mask [B,S'], labels|None, position_ids, final_ids|None).
    ``want_audio_dest``: also keep ``p.audio_dest`` (output row of every audio row) for the backward.
    ``audio_perm`` (int32, packed layout only): packed audio row r is stored at row ``audio_perm[r]`` of ``audio_rows``
    (the class order of the grouped kept-frame layout) — tasu_splice_scatter_perm."""
    _need_cuda(text_src, audio_rows, labels)
    B, S = p.B, p.S
    dev = p.input_ids.device
    H = text_src.shape[-1]
    if left_padding is None:
        left_padding = int(p.header.cpu()[L.SH_LEFT_PADDING])
    p.left_padding = int(left_padding)
    if audio_rows.dtype != text_src.dtype:
        raise TypeError("audio rows (%s) and text embeddings (%s) must share a dtype" % (audio_rows.dtype, text_src.dtype))
    if text_src.stride(-1) != 1:
        text_src = text_src.contiguous()
    if audio_rows.numel() and audio_rows.stride(-1) != 1:
        audio_rows = audio_rows.contiguous()
    if text_mode == 0:
        text2 = text_src.reshape(B * S, H)
        text_stride = text2.stride(0) if B * S > 1 else H
    else:
        text2 = text_src
        text_stride = text2.stride(0)
    if audio_layout == 1:
        audio_rows = audio_rows.contiguous()
        audio_stride = H
        n_audio_rows = p.n_audio * audio_max_len
    else:
        audio_stride = audio_rows.stride(0) if audio_rows.dim() == 2 and audio_rows.shape[0] > 1 else H
        n_audio_rows = audio_rows.shape[0] if audio_rows.dim() == 2 else 0
    # S' changes from batch to batch: allocate for S' rounded up to 32 positions so the caching allocator keeps hitting
    # the same few block sizes (a fresh size per step means cudaMalloc / cudaFree churn on a 0.3 GB tensor)
    n_pos = B * spliced_len
    cap_pos = B * ((spliced_len + 31) // 32 * 32)

    def alloc(shape_tail, dtype):
        per = 1
        for d in shape_tail:
            per *= d
        return torch.empty(max(cap_pos, 1) * per, dtype=dtype, device=dev)[:n_pos * per].view(B, spliced_len, *shape_tail)
    emb = alloc((H,), text_src.dtype)
    mask = alloc((), torch.bool if p.mask_dtype == 0 else torch.int64)
    out_labels = None
    if labels is not None:
        labels = labels.to(torch.int64).contiguous()
        out_labels = alloc((), torch.int64)
    pos = alloc((), torch.int64)
    fids = alloc((), torch.int64) if want_ids else None
    row_src = torch.empty(max(cap_pos, 1), dtype=torch.int64, device=dev)
    audio_dest = None
    if want_audio_dest:
        audio_dest = torch.empty(_cap_rows(max(n_audio_rows, 1)), dtype=torch.int32, device=dev)[:max(n_audio_rows, 1)]
        audio_dest.fill_(-1)
    L.check(L.lib().tasu_splice_scatter_perm(
        p.input_ids.data_ptr(), p.attention_mask.data_ptr(), p.mask_dtype, _ptr(labels), B, S, spliced_len, H,
        p.speech_id, text2.data_ptr(), text_mode, text_stride, audio_rows.data_ptr() if audio_rows.numel() else None,
        _ptr(audio_perm), audio_layout, audio_stride, audio_max_len, p.n_audio, _dt(emb),
        p.rowstat.data_ptr(), p.new_pos.data_ptr(), p.text_prefix.data_ptr(), p.slot_ord.data_ptr(),
        p.slot_base.data_ptr(), p.audio_off.data_ptr(), p.left_padding, pad_id, ignore_id,
        emb.data_ptr(), mask.data_ptr(), _ptr(out_labels), pos.data_ptr(), _ptr(fids), row_src.data_ptr(),
        _ptr(audio_dest), _stream()), "tasu_splice_scatter_perm")
    p.audio_dest = audio_dest
    p.row_src = row_src[:n_pos] if want_audio_dest else None      # kept for the backward of the text rows
    _count(1)
    return emb, mask, out_labels, pos, fids


def splice_audio_grad(p: SplicePlan, grad_emb: torch.Tensor, audio_layout: int, audio_max_len: int,
                      n_rows: int):
    """grad wrt the audio rows = gather of grad_emb at the audio slots (``p.audio_dest`` of the forward scatter)."""
    _need_cuda(grad_emb)
    grad_emb = grad_emb.contiguous()
    B, Sp, H = grad_emb.shape
    if p.audio_dest is None:
        raise L.TasuError("splice_audio_grad needs the forward scatter to have run with want_audio_dest=True")
    rows = p.n_audio * audio_max_len if audio_layout == 1 else n_rows
    ga = torch.empty(_cap_rows(max(rows, 1)), H, dtype=grad_emb.dtype, device=grad_emb.device)[:rows]
    L.check(L.lib().tasu_gather_rows(grad_emb.data_ptr(), _dt(grad_emb), H, p.audio_dest.data_ptr(), rows, H,
                                     ga.data_ptr(), H, _stream()), "tasu_gather_rows")
    _count(1)
    return ga.view(p.n_audio, audio_max_len, H) if audio_layout == 1 else ga


def splice_text_grad(p: SplicePlan, grad_emb: torch.Tensor) -> torch.Tensor:
    """grad wrt ``inputs_embeds [B, S, H]`` (text_mode 0): every text token receives the gradient of the output row it was
    copied to, speech / padded tokens receive zero (tasu_splice_text_grad; ps-slm.py:833-834 backward)."""
    _need_cuda(grad_emb)
    grad_emb = grad_emb.contiguous()
    B, Sp, H = grad_emb.shape
    if p.row_src is None:
        raise L.TasuError("splice_text_grad needs the forward scatter to have run with want_audio_dest=True")
    gt = torch.empty(p.B * p.S, H, dtype=grad_emb.dtype, device=grad_emb.device)
    L.check(L.lib().tasu_splice_text_grad(grad_emb.data_ptr(), _dt(grad_emb), H, p.row_src.data_ptr(), B * Sp, H, gt.data_ptr(),
                                          H, p.B * p.S, _stream()), "tasu_splice_text_grad")
    _count(1)
    return gt.view(p.B, p.S, H)


# ----------------------------------------------------------------------------- training helpers
def transpose_cast(src: torch.Tensor, rows: int, cols: int, row_scale: Optional[torch.Tensor] = None):
    """[rows, cols] (pitch = stride(0)) → bf16 [cols, pad8(rows)] = (row_scale ⊙ src)^T."""
    _need_cuda(src, row_scale)
    ld = pad_to(rows, 8)
    dst = torch.empty(cols, ld, dtype=torch.bfloat16, device=src.device)
    sstride = src.stride(0) if src.shape[0] > 1 else max(cols, src.shape[-1])
    L.check(L.lib().tasu_transpose_cast(src.data_ptr(), _dt(src), rows, cols, sstride, _ptr(row_scale),
                                        dst.data_ptr(), ld, _stream()), "tasu_transpose_cast")
    _count(1)
    return dst


def silu_fwd(z: torch.Tensor):
    _need_cuda(z)
    h = torch.empty(z.shape, dtype=torch.bfloat16, device=z.device)
    L.check(L.lib().tasu_silu_fwd(z.data_ptr(), z.numel(), h.data_ptr(), _stream()), "tasu_silu_fwd")
    _count(1)
    return h


def silu_bwd(dh: torch.Tensor, z: torch.Tensor, rstd: Optional[torch.Tensor], mean: Optional[torch.Tensor]):
    """→ (dzsT bf16 [Hb, pad8(N)], db1 [Hb], g0 [Hb])."""
    _need_cuda(dh, z)
    N, Hb = z.shape
    ld = pad_to(N, 8)
    dzsT = torch.empty(Hb, ld, dtype=torch.bfloat16, device=z.device)
    db1 = torch.empty(Hb, dtype=torch.float32, device=z.device)
    g0 = torch.empty(Hb, dtype=torch.float32, device=z.device)
    L.check(L.lib().tasu_silu_bwd(dh.data_ptr(), z.data_ptr(), N, Hb, _ptr(rstd), _ptr(mean), dzsT.data_ptr(), ld,
                                  db1.data_ptr(), g0.data_ptr(), _stream()), "tasu_silu_bwd")
    _count(1)
    return dzsT, db1, g0


def flat_scale_cast(src: torch.Tensor, dst: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    """dst = cast(scale * src) over flat contiguous buffers of equal length (tasu_flat_scale_cast)."""
    _need_cuda(src, dst)
    if src.numel() != dst.numel() or not src.is_contiguous() or not dst.is_contiguous():
        raise ValueError("flat_scale_cast needs contiguous buffers of equal length")
    L.check(L.lib().tasu_flat_scale_cast(src.data_ptr(), _dt(src), dst.data_ptr(), _dt(dst), src.numel(), float(scale), _stream()),
            "tasu_flat_scale_cast")
    _count(1)
    return dst


def colsum(src: torch.Tensor):
    _need_cuda(src)
    rows, cols = src.shape
    out = torch.empty(cols, dtype=torch.float32, device=src.device)
    L.check(L.lib().tasu_colsum(src.data_ptr(), _dt(src), rows, cols, src.stride(0) if rows > 1 else cols,
                                out.data_ptr(), _stream()), "tasu_colsum")
    _count(1)
    return out


def linear_silu_wgrad_finish(G: torch.Tensor, w1: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
                             g0: torch.Tensor, db1: torch.Tensor):
    """(dW1 [Hb,V], dgamma [V], dbeta [V]) from G = dzsT·x (fp32 [Hb, ldV])."""
    Hb, V = w1.shape
    dw1 = torch.empty(Hb, V, dtype=torch.float32, device=w1.device)
    dgamma = torch.empty(V, dtype=torch.float32, device=w1.device)
    dbeta = torch.empty(V, dtype=torch.float32, device=w1.device)
    L.check(L.lib().tasu_linear_silu_wgrad_finish(G.data_ptr(), G.stride(0), w1.data_ptr(), w1.stride(0),
                                                  gamma.data_ptr(), beta.data_ptr(), g0.data_ptr(), db1.data_ptr(), Hb, V,
                                                  dw1.data_ptr(), V, dgamma.data_ptr(), dbeta.data_ptr(), _stream()),
            "tasu_linear_silu_wgrad_finish")
    _count(1)
    return dw1, dgamma, dbeta


def split_bf16x3(src: torch.Tensor, pattern: int, col_scale: Optional[torch.Tensor] = None, want_ln: bool = False,
                 want_rowsum: bool = False, ln_eps: float = 1e-5):
    """fp32 [rows, K] → bf16 [rows, 6*pad64(K)] three-term split laid out for the A (0) or B (1) side of one long-K GEMM.
    Returns (dst, K', ln_mean, ln_rstd, row_sum)."""
    _need_cuda(src, col_scale)
    if src.shape[1] > 1 and src.stride(1) != 1:
        src = src.contiguous()
    rows, K = src.shape
    Kp = pad_to(K)
    dst = torch.empty(rows, 6 * Kp, dtype=torch.bfloat16, device=src.device)
    mean = torch.empty(rows, dtype=torch.float32, device=src.device) if want_ln else None
    rstd = torch.empty(rows, dtype=torch.float32, device=src.device) if want_ln else None
    rsum = torch.empty(rows, dtype=torch.float32, device=src.device) if want_rowsum else None
    L.check(L.lib().tasu_split_bf16x3(src.data_ptr(), _dt(src), rows, K, src.stride(0) if rows > 1 else K, _ptr(col_scale),
                                      pattern, dst.data_ptr(), 6 * Kp, _ptr(mean), _ptr(rstd), float(ln_eps), _ptr(rsum),
                                      _stream()), "tasu_split_bf16x3")
    _count(1)
    return dst, 6 * Kp, mean, rstd, rsum


def gemm_fp32x3(a_split: torch.Tensor, b_split: torch.Tensor, M: int, N: int, Ksplit: int, out: torch.Tensor,
                epilogue: int = L.EPI_NONE, bias=None, row_rstd=None, row_mean=None, colsum=None, slice_k: int = 1536,
                m_dev: Optional[torch.Tensor] = None):
    """fp32-accurate GEMM on split operands (split_bf16x3): K' is cut into slices of ``slice_k`` so the tensor core's
    truncating accumulation stays short; the slices are summed with round-to-nearest fp32 adds (tasu_sum_epilogue)."""
    n_parts = (Ksplit + slice_k - 1) // slice_k
    ldp = pad_to(N, 4)
    parts = torch.empty(n_parts, M, ldp, dtype=torch.float32, device=out.device)
    for p in range(n_parts):
        k0, k1 = p * slice_k, min(Ksplit, (p + 1) * slice_k)
        gemm_bf16_tn(a_split[:, k0:k1], b_split[:, k0:k1], M, N, k1 - k0, parts[p], m_dev=m_dev)
    L.check(L.lib().tasu_sum_epilogue(parts.data_ptr(), n_parts, M * ldp, M, N, ldp, epilogue, _ptr(bias), _ptr(row_rstd),
                                      _ptr(row_mean), _ptr(colsum), out.data_ptr(), _dt(out), out.stride(0), _stream()),
            "tasu_sum_epilogue")
    _count(1)
    return out


# ----------------------------------------------------------------------------- token-row projector (text-simulated rows)
class TokenRows:
    """Row descriptors of a text-simulated posterior batch, grouped by token (tasu_host_group_tokens).
    Row r is ``base[r]·1 + (hot[r] − base[r])·onehot(tok[r])``; rows are packed utterance after utterance."""
    __slots__ = ("n_rows", "n_uniq", "V", "hot", "base", "uniq", "seg_off", "perm", "lens", "lens_host")


def group_token_rows(tok, hot, base, lens, V: int, device) -> TokenRows:
    """Host numpy descriptors (int32 tok, fp32 hot/base, per-utterance lens) → grouped device tensors with ONE
    pinned staging buffer and ONE host→device copy."""
    import ctypes

    import numpy as np
    n = int(tok.shape[0])
    tok = np.ascontiguousarray(tok, dtype=np.int32)
    B = len(lens)
    # staging layout (int32 words): uniq[n] | seg_off[n+1] | perm[n] | hot[n] | base[n] | lens[B] (as int64 → 2B words)
    o_uniq, o_seg, o_perm, o_hot, o_base = 0, n, 2 * n + 1, 3 * n + 1, 4 * n + 1
    o_lens = 5 * n + 1 + ((5 * n + 1) & 1)                       # 8-byte aligned
    words = o_lens + 2 * B
    stage = torch.empty(max(words, 2), dtype=torch.int32, pin_memory=torch.cuda.is_available())
    s = stage.numpy()
    n_uniq = ctypes.c_int32(0)
    i32p = lambda a, off=0: ctypes.c_void_p(a.ctypes.data + 4 * off)   # noqa: E731
    L.check(L.lib().tasu_host_group_tokens(i32p(tok), n, V, i32p(s, o_uniq), i32p(s, o_seg), i32p(s, o_perm),
                                           ctypes.cast(ctypes.pointer(n_uniq), ctypes.c_void_p)),
            "tasu_host_group_tokens")
    s[o_hot:o_hot + n] = np.ascontiguousarray(hot, dtype=np.float32).view(np.int32)
    s[o_base:o_base + n] = np.ascontiguousarray(base, dtype=np.float32).view(np.int32)
    s[o_lens:o_lens + 2 * B] = np.asarray(lens, dtype=np.int64).view(np.int32)
    d = stage.to(device, non_blocking=True)
    t = TokenRows()
    t.n_rows, t.n_uniq, t.V = n, int(n_uniq.value), V
    t.uniq = d[o_uniq:o_uniq + t.n_uniq]
    t.seg_off = d[o_seg:o_seg + t.n_uniq + 1]
    t.perm = d[o_perm:o_perm + n]
    t.hot = d[o_hot:o_hot + n].view(torch.float32)
    t.base = d[o_base:o_base + n].view(torch.float32)
    t.lens = d[o_lens:o_lens + 2 * B].view(torch.int64)
    t.lens_host = [int(x) for x in lens]
    return t


class TokenBatch:
    """Tokenised transcripts of one batch, flattened once (the data loader's output): int32 ids + int64 lengths."""
    __slots__ = ("tok", "lens", "B", "total")

    def __init__(self, ids_list):
        import numpy as np
        self.B = len(ids_list)
        self.lens = np.fromiter((len(i) for i in ids_list), dtype=np.int64, count=self.B)
        self.total = int(self.lens.sum())
        self.tok = (np.concatenate([np.asarray(i, dtype=np.int32) for i in ids_list]) if self.total
                    else np.zeros(0, dtype=np.int32))


def sim_token_rows(batch: TokenBatch, V: int, device, drop_prob: float = 0.05, smooth_low: float = 0.0,
                   smooth_high: float = 0.1) -> TokenRows:
    """Noisy simulator (ps-slm.py:360-409, insert_prob = 0) → grouped descriptors on the device: one ``torch.rand``
    call (the reference's RNG stream), one native host call, one host→device copy."""
    import ctypes
    cap, B = batch.total, batch.B
    u = torch.rand(cap + B)
    o_uniq, o_seg, o_perm, o_hot, o_base = 0, cap, 2 * cap + 1, 3 * cap + 1, 4 * cap + 1
    o_lens = 5 * cap + 1 + ((5 * cap + 1) & 1)
    words = o_lens + 2 * B
    stage = torch.empty(max(words, 2), dtype=torch.int32, pin_memory=torch.cuda.is_available())
    n_rows, n_uniq = ctypes.c_int64(0), ctypes.c_int32(0)
    L.check(L.lib().tasu_host_sim_token_rows(u.data_ptr(), batch.tok.ctypes.data, batch.lens.ctypes.data, B, V,
                                             float(drop_prob), float(smooth_low), float(smooth_high), stage.data_ptr(),
                                             words, ctypes.addressof(n_rows), ctypes.addressof(n_uniq)),
            "tasu_host_sim_token_rows")
    n = int(n_rows.value)
    d = stage.to(device, non_blocking=True)
    t = TokenRows()
    t.n_rows, t.n_uniq, t.V = n, int(n_uniq.value), V
    t.uniq = d[o_uniq:o_uniq + t.n_uniq]
    t.seg_off = d[o_seg:o_seg + t.n_uniq + 1]
    t.perm = d[o_perm:o_perm + n]
    t.hot = d[o_hot:o_hot + n].view(torch.float32)
    t.base = d[o_base:o_base + n].view(torch.float32)
    t.lens = d[o_lens:o_lens + 2 * B].view(torch.int64)
    t.lens_host = stage[o_lens:o_lens + 2 * B].view(torch.int64).tolist()
    return t


def linear_rowdots(w1: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, b1: Optional[torch.Tensor]):
    """(S = W1·γ, D = W1·β + b1), fp32 [N] each."""
    _need_cuda(w1, gamma, beta, b1)
    N, K = w1.shape
    w1 = w1.float().contiguous()
    S = torch.empty(N, dtype=torch.float32, device=w1.device)
    D = torch.empty(N, dtype=torch.float32, device=w1.device)
    L.check(L.lib().tasu_linear_rowdots(w1.data_ptr(), w1.stride(0), gamma.float().contiguous().data_ptr(),
                                        beta.float().contiguous().data_ptr(),
                                        _ptr(b1.float().contiguous()) if b1 is not None else None, N, K,
                                        S.data_ptr(), D.data_ptr(), _stream()), "tasu_linear_rowdots")
    _count(1)
    return S, D


def tokrow_cols(w1: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, b1: Optional[torch.Tensor], rows: TokenRows):
    """One dense pass over W1 → (colT fp32 [n_uniq, Hb], S, D)."""
    _need_cuda(w1, gamma, beta, b1)
    Hb, V = w1.shape
    dev = w1.device
    w1 = w1.float().contiguous()
    colT = torch.empty(max(rows.n_uniq, 1), Hb, dtype=torch.float32, device=dev)
    S = torch.empty(Hb, dtype=torch.float32, device=dev)
    D = torch.empty(Hb, dtype=torch.float32, device=dev)
    nbytes = L.lib().tasu_tokrow_cols_workspace(V, Hb)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    L.check(L.lib().tasu_tokrow_cols(w1.data_ptr(), w1.stride(0), gamma.float().contiguous().data_ptr(),
                                     beta.float().contiguous().data_ptr(),
                                     _ptr(b1.float().contiguous()) if b1 is not None else None, rows.uniq.data_ptr(),
                                     rows.n_uniq, V, Hb, colT.data_ptr(), S.data_ptr(), D.data_ptr(), ws.data_ptr(),
                                     nbytes, _stream()), "tasu_tokrow_cols")
    _count(3)
    return colT, S, D


def tokrow_fwd(w1: torch.Tensor, gamma: torch.Tensor, S: torch.Tensor, D: torch.Tensor, rows: TokenRows,
               ln_eps: float = 1e-5, want_z: bool = True, colT: Optional[torch.Tensor] = None):
    """→ (z fp32 [n, Hb] | None, h bf16 [n, Hb], row_a, row_e)."""
    _need_cuda(w1, gamma, S, D)
    Hb, V = w1.shape
    n = rows.n_rows
    dev = w1.device
    w1 = w1.float().contiguous()
    z = torch.empty(max(n, 1), Hb, dtype=torch.float32, device=dev)[:n] if want_z else None
    h = torch.empty(max(n, 1), Hb, dtype=torch.bfloat16, device=dev)[:n]
    row_a = torch.empty(max(n, 1), dtype=torch.float32, device=dev)
    row_e = torch.empty(max(n, 1), dtype=torch.float32, device=dev)
    if colT is not None:                                       # training forward: warp-per-row pass on compact columns
        slot_ws = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        L.check(L.lib().tasu_tokrow_rows_fwd(colT.data_ptr(), S.data_ptr(), D.data_ptr(), rows.seg_off.data_ptr(),
                                             rows.perm.data_ptr(), rows.hot.data_ptr(), rows.base.data_ptr(), rows.n_uniq, n,
                                             V, Hb, float(ln_eps), _ptr(z), h.data_ptr(), row_a.data_ptr(), row_e.data_ptr(),
                                             slot_ws.data_ptr(), _stream()), "tasu_tokrow_rows_fwd")
        _count(2)
        return z, h, row_a, row_e
    L.check(L.lib().tasu_tokrow_fwd(w1.data_ptr(), w1.stride(0), gamma.float().contiguous().data_ptr(), S.data_ptr(),
                                    D.data_ptr(), rows.uniq.data_ptr(), rows.seg_off.data_ptr(), rows.perm.data_ptr(),
                                    rows.hot.data_ptr(), rows.base.data_ptr(), rows.n_uniq, n, V, Hb, float(ln_eps),
                                    _ptr(z), h.data_ptr(), row_a.data_ptr(), row_e.data_ptr(), _ptr(colT), _stream()),
            "tasu_tokrow_fwd")
    _count(1)
    return z, h, row_a, row_e


def tokrow_bwd_rows(dh: torch.Tensor, z: torch.Tensor, rows: TokenRows, row_a: torch.Tensor, row_e: torch.Tensor):
    """→ (P fp32 [n_uniq, Hb], db1 [Hb], E [Hb])."""
    _need_cuda(dh, z)
    n, Hb = z.shape
    dev = z.device
    P = torch.empty(max(rows.n_uniq, 1), Hb, dtype=torch.float32, device=dev)
    db1 = torch.empty(Hb, dtype=torch.float32, device=dev)
    E = torch.empty(Hb, dtype=torch.float32, device=dev)
    nbytes = L.lib().tasu_tokrow_bwd_workspace(Hb, rows.n_uniq)
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
    L.check(L.lib().tasu_tokrow_bwd_rows(dh.data_ptr(), z.data_ptr(), n, Hb, rows.seg_off.data_ptr(), rows.perm.data_ptr(),
                                         row_a.data_ptr(), row_e.data_ptr(), rows.n_uniq, P.data_ptr(), db1.data_ptr(),
                                         E.data_ptr(), ws.data_ptr(), nbytes, _stream()), "tasu_tokrow_bwd_rows")
    _count(2)
    return P, db1, E


def tokrow_wgrad_finish(P: torch.Tensor, rows: TokenRows, w1: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
                        E: torch.Tensor, db1: torch.Tensor):
    """(dW1 [Hb, V], dgamma [V], dbeta [V]) in one pass over W1."""
    Hb, V = w1.shape
    dev = w1.device
    w1 = w1.float().contiguous()
    dw1 = torch.empty(Hb, V, dtype=torch.float32, device=dev)
    dgamma = torch.empty(V, dtype=torch.float32, device=dev)
    dbeta = torch.empty(V, dtype=torch.float32, device=dev)
    slot = torch.empty(V, dtype=torch.int32, device=dev)
    L.check(L.lib().tasu_tokrow_wgrad_finish(P.data_ptr(), rows.uniq.data_ptr(), rows.n_uniq, slot.data_ptr(), w1.data_ptr(),
                                             w1.stride(0), gamma.float().contiguous().data_ptr(),
                                             beta.float().contiguous().data_ptr(), E.data_ptr(), db1.data_ptr(), Hb, V,
                                             dw1.data_ptr(), V, dgamma.data_ptr(), dbeta.data_ptr(), _stream()),
            "tasu_tokrow_wgrad_finish")
    _count(2)
    return dw1, dgamma, dbeta


def _f32c(t: torch.Tensor) -> torch.Tensor:
    t = t.detach()
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()


_TOKROW_WS = {}     # device → grow-only scratch buffer (stream-ordered reuse; holds no state between calls)


def _cap_rows(n: int, q: int = 2048) -> int:
    """Data-dependent row counts are rounded up for allocation so the caching allocator sees a few distinct sizes."""
    return max(q, (n + q - 1) // q * q)


def tokrow_train_workspace(rows: TokenRows, Hb: int, H: int, device) -> torch.Tensor:
    nbytes = L.lib().tasu_tokrow_train_workspace(_cap_rows(rows.n_rows), _cap_rows(rows.n_uniq), rows.V, Hb, H)
    key = (torch.device(device).index, torch.cuda.current_stream().cuda_stream)
    ws = _TOKROW_WS.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = None
        _TOKROW_WS[key] = None
        ws = torch.empty(int(nbytes * 1.25), dtype=torch.uint8, device=device)
        _TOKROW_WS[key] = ws
    return ws


def tokrow_linear_silu_fwd(rows: TokenRows, gamma, beta, w1, b1, w2, b2, eps: float, out_dtype, ws: torch.Tensor,
                           want_z: bool = True):
    """Composite forward (tasu_tokrow_linear_silu_fwd) → (y [n, H], z fp32 [n, Hb] | None, h bf16, row_a, row_e)."""
    _need_cuda(w1, w2)
    Hb, V = w1.shape
    H = w2.shape[0]
    n, dev = rows.n_rows, w1.device
    w1, w2, gamma, beta, b1, b2 = _f32c(w1), _f32c(w2), _f32c(gamma), _f32c(beta), _f32c(b1), _f32c(b2)
    cap = _cap_rows(n)
    z = torch.empty(cap, Hb, dtype=torch.float32, device=dev)[:n] if want_z else None
    h = torch.empty(cap, Hb, dtype=torch.bfloat16, device=dev)[:n]
    ab = torch.empty(2, cap, dtype=torch.float32, device=dev)
    y = torch.empty(cap, H, dtype=out_dtype, device=dev)[:n]
    L.check(L.lib().tasu_tokrow_linear_silu_fwd(
        w1.data_ptr(), w1.stride(0), gamma.data_ptr(), beta.data_ptr(), b1.data_ptr(), w2.data_ptr(), w2.stride(0),
        b2.data_ptr(), rows.uniq.data_ptr(), rows.seg_off.data_ptr(), rows.perm.data_ptr(), rows.hot.data_ptr(),
        rows.base.data_ptr(), rows.n_uniq, n, V, Hb, H, float(eps), _ptr(z), h.data_ptr(), ab[0].data_ptr(),
        ab[1].data_ptr(), y.data_ptr(), _dt(y), H, ws.data_ptr(), ws.numel(), _stream()), "tasu_tokrow_linear_silu_fwd")
    _count(4)
    return y, z, h, ab[0], ab[1]


def tokrow_linear_silu_bwd(dy: torch.Tensor, rows: TokenRows, z, h, row_a, row_e, gamma, beta, w1, w2, ws: torch.Tensor,
                           between=None):
    """Composite backward (tasu_tokrow_linear_silu_bwd) → (dgamma, dbeta, dW1, db1, dW2, db2), all fp32 views of ONE
    flat buffer in parameter order.  ``between(flat, n_first)``: optional callback invoked after the W1 half
    (flat[:n_first] = dgamma | dbeta | dW1 | db1 is enqueued) and before the W2 half — the hook for overlapping a
    gradient all-reduce with the rest of the backward."""
    Hb, V = w1.shape
    H = w2.shape[0]
    n, dev = rows.n_rows, w1.device
    w1, w2, gamma, beta = _f32c(w1), _f32c(w2), _f32c(gamma), _f32c(beta)
    # ONE flat gradient buffer in parameter order (norm.weight, norm.bias, ffn.0.weight, ffn.0.bias, ffn.2.weight,
    # ffn.2.bias), every slice 256-byte aligned: the gradients autograd hands to the parameters are views of it, so
    # the data-parallel all-reduce (dist.allreduce_gradients) runs in place on one message without a flatten/copy
    sizes = (V, V, Hb * V, Hb, H * Hb, H)
    offs, o = [], 0
    for n_el in sizes:
        offs.append(o)
        o += pad_to(n_el)
    flat = torch.empty(o, dtype=torch.float32, device=dev)
    dgamma, dbeta = flat[offs[0]:offs[0] + V], flat[offs[1]:offs[1] + V]
    dw1 = flat[offs[2]:offs[2] + Hb * V].view(Hb, V)
    db1 = flat[offs[3]:offs[3] + Hb]
    dw2 = flat[offs[4]:offs[4] + H * Hb].view(H, Hb)
    db2 = flat[offs[5]:offs[5] + H]

    def call(phase):
        L.check(L.lib().tasu_tokrow_linear_silu_bwd(
            dy.data_ptr(), _dt(dy), dy.stride(0) if n > 1 else H, _ptr(z), h.data_ptr(), row_a.data_ptr(), row_e.data_ptr(),
            w1.data_ptr(), w1.stride(0), gamma.data_ptr(), beta.data_ptr(), w2.data_ptr(), w2.stride(0),
            rows.uniq.data_ptr(), rows.seg_off.data_ptr(), rows.perm.data_ptr(), rows.n_uniq, n, V, Hb, H,
            dw1.data_ptr(), V, dgamma.data_ptr(), dbeta.data_ptr(), db1.data_ptr(), dw2.data_ptr(), Hb, db2.data_ptr(),
            phase, ws.data_ptr(), ws.numel(), _stream()), "tasu_tokrow_linear_silu_bwd")
    if between is None:
        call(0)
    else:
        call(1)
        between(flat, offs[4])
        call(2)
    _count(10)
    return dgamma, dbeta, dw1, db1, dw2, db2
