"""Host side of the bridge: the reference's method-level API on top of the C ABI.

Function names, argument meaning, return tuples and error behaviour mirror
``slam_model_asr`` in Multitask/model/ps-slm.py (paths relative to /root/reference/):

* ``psd``                                   ← ps-slm.py:237-317
* ``merge_input_ids_with_audio_features``   ← ps-slm.py:679-873
* ``ctc_head_psd_project_splice`` (class ``TasuBridge``) ← the dispatch of
  ps-slm.py:581-658 with the shipped inference flags (ctc_posterior, do_psd,
  linear-silu projector), fused so that the [B,T,25055] posterior is never
  materialised and the host reads back exactly one 96-byte header per batch.

Everything here is tensor plumbing; the arithmetic lives in libtasu_bridge.so.
"""
import os
from typing import Optional, Tuple

import torch

from . import _lib as L
from . import ops

BLANK_THRESHOLD = 0.90


def _as_compute(x: torch.Tensor) -> torch.Tensor:
    if x.dtype in (torch.float32, torch.bfloat16):
        return x
    return x.float()


def psd(encoder_out: torch.Tensor, encoder_out_lens: torch.Tensor, ctc_posterior: torch.Tensor,
        blank_id: int = 0, blank_threshold: float = BLANK_THRESHOLD) -> Tuple[torch.Tensor, torch.Tensor]:
    """Drop-in for ``slam_model_asr.psd`` (ps-slm.py:237-317).

    Returns ``(encoder_outs [B, max_b M_b, D], new_lens [B] int64 on device)``; padded rows
    are zeros; all-empty input gives ``[B, 0, D]`` (ps-slm.py:304-306).  One device→host
    read (the 32-byte collapse header) instead of the reference's O(B·T) syncs."""
    encoder_out = _as_compute(encoder_out)
    ctc_posterior = _as_compute(ctc_posterior)
    B, T, D = encoder_out.shape
    if ctc_posterior.shape[0] != B or ctc_posterior.shape[1] != T:
        raise ValueError("encoder_out and ctc_posterior must agree on [B, T]")
    lens = encoder_out_lens.to(device=encoder_out.device, dtype=torch.int64)
    st = ops.frame_stats(ctc_posterior, L.INPUT_PROBS, blank_id)
    plan = ops.collapse_plan(st, lens, blank_id, blank_threshold)
    hdr = plan.header.cpu()
    max_len = int(hdr[L.CH_MAX_LEN])
    if max_len == 0:
        return encoder_out.new_zeros(B, 0, D), torch.zeros(B, dtype=torch.long, device=encoder_out.device)
    out = torch.empty(B, max_len, D, dtype=encoder_out.dtype, device=encoder_out.device)
    ops.segment_meanpool(encoder_out, plan, 1, max_len, B * max_len, out, D)
    return out, plan.new_lens


def psd_from_encoder(raw_encoder_out: torch.Tensor, raw_encoder_out_lens: torch.Tensor, w_ctc_bf16: torch.Tensor,
                     b_ctc: torch.Tensor, blank_id: int = 0, blank_threshold: float = BLANK_THRESHOLD, n_prefix: int = 4):
    """Raw-feature branch (``ctc_posterior=False, do_psd=True``; ps-slm.py:450-454 + :515-518 / :581-585 + :645-648):
    ``psd(encoder_out, lens, softmax(ctc_lo(raw))[:, 4:])`` WITHOUT materialising the ``[B, T, 25055]`` posterior —
    the segmentation comes from the fused CTC-head statistics (argmax / blank probability out of the tcgen05 GEMM
    epilogue), the 512-d encoder frames themselves are mean-pooled.  Same return contract as ``psd``."""
    B, T4, Denc = raw_encoder_out.shape
    T = T4 - n_prefix
    V = w_ctc_bf16.shape[0]
    dev = raw_encoder_out.device
    feats = _as_compute(raw_encoder_out)
    x2 = feats.reshape(B * T4, Denc)
    xb = x2 if x2.dtype == torch.bfloat16 and Denc % 64 == 0 else ops.cast_rows(x2, torch.bfloat16, ops.pad_to(Denc))[0]
    lens = torch.clamp(raw_encoder_out_lens.to(device=dev, dtype=torch.int64) - n_prefix, min=0)
    st = ops.ctc_head_stats(xb, w_ctc_bf16, b_ctc, B, T, n_prefix, V, Denc, blank_id)
    plan = ops.collapse_plan(st, lens, blank_id, blank_threshold)
    max_len = int(plan.header.cpu()[L.CH_MAX_LEN])
    if max_len == 0:
        return feats.new_zeros(B, 0, Denc), torch.zeros(B, dtype=torch.long, device=dev)
    out = torch.empty(B, max_len, Denc, dtype=feats.dtype, device=dev)
    ops.segment_meanpool(feats[:, n_prefix:, :], plan, 1, max_len, B * max_len, out, Denc)
    return out, plan.new_lens


# voca_trans branch without the [B, T', V_llm] logits tensor (TASU_VOCA_FUSED=0: materialised logits, round-1 path)
FUSED_VOCA_TRANS = os.environ.get("TASU_VOCA_FUSED", "1") != "0"
LLM_BLANK_ID = 151643     # blank index of the LLM-vocabulary CTC head, hard-coded in the reference (ps-slm.py:491, :621)


class _VocaTransFunction(torch.autograd.Function):
    """out = softmax(pooled logits without the blank column) · E[:V_real], logits = CTC head over the LLM vocabulary
    (``map``) applied to the k-concatenated encoder frames, PSD mean-pooling of the LOGITS in between.  Gradient to the
    head only:  dProbs = dOut·Eᵀ,  dS = P∘(dProbs − dOut·Out) (tasu_attn_score_grad),  and — mean-pooling being linear —
    dW = dSᵀ·x̄ with x̄ the SAME segmented mean of the input frames, db = Σ_rows dS; no un-pooling is needed."""

    @staticmethod
    def forward(ctx, w_map, b_map, projector, encoder_out, lens, table, do_psd, top1_emb, blank_id, threshold):
        if FUSED_VOCA_TRANS and (not do_psd or blank_id == w_map.shape[0] - 1):
            return _VocaTransFunction._forward_fused(ctx, w_map, b_map, projector, encoder_out, lens, table, do_psd,
                                                     top1_emb, blank_id, threshold)
        with torch.no_grad():
            logits = projector(encoder_out)                                    # [B, T', Vp] (view of a padded buffer)
        B, Tp, Vp = logits.shape
        dev = logits.device
        H = table.shape[1]
        k = projector.k
        feat_len = lens.to(device=dev, dtype=torch.int64) // k
        x = encoder_out
        if x.shape[1] % k:
            x = x[:, :x.shape[1] - x.shape[1] % k]
        xk = x.contiguous().view(B, Tp, -1)                                    # the k-concat the head saw (projector.py:19-24)
        xk = xk if xk.dtype in (torch.float32, torch.bfloat16) else xk.float()
        if do_psd:
            st = ops.frame_stats(logits, L.INPUT_LOGITS, blank_id, feat_len)
            plan = ops.collapse_plan(st, feat_len, blank_id, threshold)
            max_len = int(plan.header.cpu()[L.CH_MAX_LEN])
            if max_len == 0:
                ctx.empty = True
                ctx.shapes = (w_map.shape, b_map.shape)
                return torch.zeros(B, 0, H, dtype=torch.float32, device=dev), torch.zeros(B, dtype=torch.long, device=dev)
            pitch = ops.pad_to(Vp, 4)
            rows = torch.empty(B * max_len, pitch, dtype=logits.dtype, device=dev)
            ops.segment_meanpool(logits, plan, 1, max_len, B * max_len, rows, pitch)      # mean of the LOGITS (ps-slm.py:286)
            xbar = torch.empty(B * max_len, xk.shape[2], dtype=xk.dtype, device=dev)
            ops.segment_meanpool(xk, plan, 1, max_len, B * max_len, xbar, xk.shape[2])    # same segments, input side
            new_lens, V_real = plan.new_lens, Vp - 1                                      # :493 drops the blank column
        else:
            rows = logits.reshape(B * Tp, Vp) if logits.is_contiguous() else logits.contiguous().view(B * Tp, Vp)
            xbar = xk.reshape(B * Tp, -1)
            new_lens, max_len, V_real = feat_len, Tp, Vp                                  # :509-511 keeps every column
        N = rows.shape[0]
        st2 = ops.frame_stats(rows.unsqueeze(0)[:, :, :V_real], L.INPUT_LOGITS, 0)
        ctx.empty = False
        if top1_emb:                                                                      # :498-502, :512-516 (no gradient)
            out = ops.gather_rows(table, st2.argmax)
            ctx.top1 = True
            ctx.shapes = (w_map.shape, b_map.shape)
        else:
            probs = ops.softmax_rows(rows, V_real, st2)
            out = torch.empty(N, H, dtype=torch.float32, device=dev)
            ops.gemm_bf16_f32(probs, False, table[:V_real], True, N, H, V_real, out)
            ctx.top1 = False
            ctx.save_for_backward(probs, out, xbar, table)
            ctx.dims = (N, H, V_real, Vp, xbar.shape[1])
        ctx.mark_non_differentiable(new_lens)
        return out.view(B, max_len, H), new_lens

    @staticmethod
    def _forward_fused(ctx, w_map, b_map, projector, encoder_out, lens, table, do_psd, top1_emb, blank_id, threshold):
        """The same branch without the ``[B, T', Vp]`` logits (9.7 GB of fp32 at B = 64 x 30 s): the head is linear, so the
        mean of a run's logits is the logits of the run's mean input frame.  (1) greedy statistics of every frame straight
        out of the head GEMM's epilogue (``tasu_ctc_head_stats``, as on the main path) → collapse plan; (2) segmented mean
        of the INPUT frames (the x̄ the backward needs anyway); (3) statistics and bf16 probabilities of the pooled rows
        from two more passes of the head over x̄ only (blank column dropped: it is the last class), (4) ``P·E`` with the
        table read in place.  Rows beyond an utterance's compressed length are what the reference's zero-padded logits
        give: the uniform mixture of the table rows (``top1_emb``: row 0)."""
        k = projector.k
        x = encoder_out
        if x.shape[1] % k:
            x = x[:, :x.shape[1] - x.shape[1] % k]
        B = x.shape[0]
        Tp = x.shape[1] // k
        xk = x.contiguous().view(B, Tp, -1)                                    # the k-concat the head sees (projector.py:19-24)
        xk = xk if xk.dtype in (torch.float32, torch.bfloat16) else xk.float()
        Dk, Vp, H = xk.shape[2], w_map.shape[0], table.shape[1]
        dev = xk.device
        feat_len = lens.to(device=dev, dtype=torch.int64) // k
        with torch.no_grad():
            w, b = projector.head_bf16()
        ldk = ops.pad_to(Dk)
        if do_psd:
            xb = ops.cast_rows(xk.reshape(B * Tp, Dk), torch.bfloat16, ldk)[0]
            st = ops.ctc_head_stats(xb, w, b, B, Tp, 0, Vp, Dk, blank_id)
            plan = ops.collapse_plan(st, feat_len, blank_id, threshold)
            max_len = int(plan.header.cpu()[L.CH_MAX_LEN])
            if max_len == 0:
                ctx.empty = True
                ctx.shapes = (w_map.shape, b_map.shape)
                return torch.zeros(B, 0, H, dtype=torch.float32, device=dev), torch.zeros(B, dtype=torch.long, device=dev)
            xbar = torch.empty(B * max_len, Dk, dtype=xk.dtype, device=dev)
            ops.segment_meanpool(xk, plan, 1, max_len, B * max_len, xbar, Dk)  # zero rows beyond the compressed lengths
            new_lens, V_real = plan.new_lens, Vp - 1                           # :493 drops the blank column
        else:
            xbar = xk.reshape(B * Tp, Dk)
            new_lens, max_len, V_real = feat_len, Tp, Vp                       # :509-511 keeps every column
        N = xbar.shape[0]
        xbb = ops.cast_rows(xbar, torch.bfloat16, ldk)[0]
        st2 = ops.ctc_head_stats(xbb, w[:V_real], b[:V_real], 1, N, 0, V_real, Dk, 0)
        pad = (torch.arange(max_len, device=dev)[None, :] >= new_lens[:, None]).reshape(N) if do_psd else None
        ctx.empty = False
        if top1_emb:                                                           # :498-502, :512-516 (no gradient)
            idx = st2.argmax if pad is None else st2.argmax.masked_fill(pad, 0)
            out = ops.gather_rows(table, idx)
            ctx.top1 = True
            ctx.shapes = (w_map.shape, b_map.shape)
        else:
            probs = torch.empty(N, ops.pad_to(V_real), dtype=torch.bfloat16, device=dev)
            ops.gemm_bf16_tn(xbb, w[:V_real], N, V_real, Dk, probs, L.EPI_SOFTMAX, b[:V_real],
                             torch.reciprocal(st2.row_sumexp), st2.row_max)
            out = torch.empty(N, H, dtype=torch.float32, device=dev)
            ops.gemm_bf16_f32(probs, False, table[:V_real], True, N, H, V_real, out)
            if pad is not None:
                uniform = table[:V_real].mean(0, dtype=torch.float32)          # softmax of an all-zero logits row · E
                out = torch.where(pad[:, None], uniform[None, :], out)
            ctx.top1 = False
            ctx.pad = pad                                                      # padding rows carry no gradient (backward)
            ctx.save_for_backward(probs, out, xbar, table)
            ctx.dims = (N, H, V_real, Vp, Dk)
        ctx.mark_non_differentiable(new_lens)
        return out.view(B, max_len, H), new_lens

    @staticmethod
    def backward(ctx, dout, _):
        if ctx.empty or ctx.top1:
            ws, bs = ctx.shapes
            return (torch.zeros(ws, device=dout.device), torch.zeros(bs, device=dout.device)) + (None,) * 8
        probs, out, xbar, table = ctx.saved_tensors
        N, H, V_real, Vp, Dk = ctx.dims
        dev = dout.device
        do = dout.reshape(N, H).float().contiguous()
        pad = getattr(ctx, "pad", None)
        if pad is not None:
            # rows beyond the compressed lengths are constants of the forward (zero logits in the reference): with a zero
            # upstream gradient dProbs and dS of those rows vanish
            do = do.masked_fill(pad[:, None], 0.0)
        dob, _, _ = ops.cast_rows(do, torch.bfloat16, ops.pad_to(H, 8))
        dP = torch.empty(N, ops.pad_to(V_real, 4), dtype=torch.float32, device=dev)
        ops.gemm_bf16_tn(dob, table, N, V_real, H, dP)                                    # dProbs = dOut · Eᵀ
        ldv = probs.stride(0)
        dS = torch.empty(N, ldv, dtype=torch.bfloat16, device=dev)
        L.check(L.lib().tasu_attn_score_grad(probs.data_ptr(), ldv, dP.data_ptr(), dP.stride(0), do.data_ptr(),
                                             out.data_ptr(), H, H, N, V_real, dS.data_ptr(), ldv, ops._stream()),
                "tasu_attn_score_grad")
        ops._count(1)
        xb, _, _ = ops.cast_rows(xbar if xbar.dim() == 2 else xbar.reshape(N, Dk), torch.bfloat16, ops.pad_to(Dk, 8))
        dw = torch.zeros(Vp, ops.pad_to(Dk, 4), dtype=torch.float32, device=dev)          # blank row (do_psd) stays zero
        ops.gemm_bf16_f32(dS[:, :V_real], True, xb[:, :Dk], True, V_real, Dk, N, dw[:V_real, :Dk])   # dW = dSᵀ · x̄
        db = torch.zeros(Vp, dtype=torch.float32, device=dev)
        db[:V_real] = ops.colsum(dS[:, :V_real])
        return (dw[:, :Dk], db) + (None,) * 8


def voca_trans_project(projector, encoder_out: torch.Tensor, encoder_out_lens: torch.Tensor, embed_table_bf16: torch.Tensor,
                       do_psd: bool, top1_emb: bool, blank_id: int = LLM_BLANK_ID, blank_threshold: float = BLANK_THRESHOLD):
    """Vocabulary-transfer branch (``voca_trans=True``; ps-slm.py:485-513 / :615-643) with the input the branch evidently
    means — the reference reads ``encoder_outs`` there before assigning it (UnboundLocalError as shipped); the only
    tensor of the right shape in scope is ``encoder_out``:

      logits = simple_linear projector (a CTC head over the LLM vocabulary, projector.py:10-26)
      do_psd: psd(features = logits, posterior = softmax(logits), blank = 151643)  → pooled logits, blank column dropped
      softmax over the remaining columns · embed_matrix[:V_real]   (or the embedding row of the top-1 id, ``top1_emb``)

    No ``[B, T, V]`` softmax tensor is materialised: greedy statistics come from ``tasu_frame_stats`` on the logits,
    the contraction reads the embedding table in place (MN-major B operand of ``tasu_gemm_bf16_f32``).  Trainable: the
    head (``map.weight`` / ``map.bias``) receives its gradient through ``_VocaTransFunction``.
    Returns ``(projector_outs [B, T_new, H] fp32 | table dtype for top1, lengths [B] int64)``."""
    if encoder_out.requires_grad:
        raise NotImplementedError("the bridge does not propagate a gradient to the encoder output")
    return _VocaTransFunction.apply(projector.map.weight, projector.map.bias, projector, encoder_out, encoder_out_lens,
                                    embed_table_bf16, bool(do_psd), bool(top1_emb), int(blank_id), float(blank_threshold))


def merge_input_ids_with_audio_features(audio_features: torch.Tensor, num_audio_tokens: torch.Tensor,
                                        inputs_embeds: torch.Tensor, input_ids: torch.Tensor,
                                        attention_mask: torch.Tensor, labels: Optional[torch.Tensor],
                                        speech_id: int, pad_id: int, ignore_id: int = -100):
    """Drop-in for ``slam_model_asr._merge_input_ids_with_audio_features`` (ps-slm.py:679-873).

    Same 5-tuple ``(final_embedding, final_attention_mask, final_labels|None, position_ids,
    final_input_ids)`` and the same two ``ValueError``s (:783-785, :861-865)."""
    if audio_features.dim() != 3:
        raise ValueError("audio_features must be [num_audios, max_audio_tokens, embed_dim]")
    if audio_features.dtype != inputs_embeds.dtype:
        audio_features = audio_features.to(inputs_embeds.dtype)
    p = ops.splice_rowstat(input_ids, attention_mask, speech_id)
    ops.splice_plan(p, num_audio_tokens, 1)
    hdr = p.header.cpu()
    _raise_splice_errors(hdr, attention_mask, num_audio_tokens.numel())
    if torch.is_grad_enabled() and (audio_features.requires_grad or inputs_embeds.requires_grad):
        from .autograd import SpliceFunction            # gradient flows back to the projector output and the text embeddings
        p.left_padding = int(hdr[L.SH_LEFT_PADDING])
        return SpliceFunction.apply(audio_features, inputs_embeds, p, int(hdr[L.SH_SPLICED_LEN]), 0, 1,
                                    audio_features.shape[1], labels, pad_id, ignore_id)
    return ops.splice_scatter(p, int(hdr[L.SH_SPLICED_LEN]), inputs_embeds, 0, audio_features, 1,
                              audio_features.shape[1], labels, pad_id, ignore_id,
                              left_padding=int(hdr[L.SH_LEFT_PADDING]))


_PINNED_HDR = []


def _pinned_header():
    """Small ring of pinned host header buffers (UVA-mapped: the plan kernels write them directly)."""
    if not _PINNED_HDR:
        _PINNED_HDR.extend([[torch.zeros(L.SH_WORDS, dtype=torch.int64).pin_memory() for _ in range(8)], 0])
    ring = _PINNED_HDR[0]
    _PINNED_HDR[1] = (_PINNED_HDR[1] + 1) % len(ring)
    return ring[_PINNED_HDR[1]]


class PendingSplicePlan:
    """Splice plan whose header lands in pinned host memory; ``finish()`` waits for the plan kernels only, so work
    enqueued after ``begin_splice_plan`` (e.g. the projector) keeps the GPU busy while the host reads S'."""
    __slots__ = ("plan", "header", "event", "attention_mask", "n_audio")

    def finish(self):
        self.event.synchronize()
        hdr = self.header.clone()
        _raise_splice_errors(hdr, self.attention_mask, self.n_audio)
        self.plan.left_padding = int(hdr[L.SH_LEFT_PADDING])
        return self.plan, int(hdr[L.SH_SPLICED_LEN])


def begin_splice_plan(input_ids: torch.Tensor, attention_mask: torch.Tensor, num_audio_tokens: torch.Tensor,
                      speech_id: int, div_k: int = 1) -> PendingSplicePlan:
    """Integer part of step 4 (ps-slm.py:765-812, :842-865), enqueued early."""
    pend = PendingSplicePlan()
    pend.header = _pinned_header()
    pend.plan = ops.splice_rowstat(input_ids, attention_mask, speech_id)
    ops.splice_plan(pend.plan, num_audio_tokens, div_k, header=pend.header)
    pend.event = torch.cuda.Event()
    pend.event.record()
    pend.attention_mask, pend.n_audio = attention_mask, num_audio_tokens.numel()
    return pend


def merge_packed_audio_rows(audio_rows: torch.Tensor, num_audio_tokens: torch.Tensor, max_audio_tokens: int,
                            text_src: torch.Tensor, text_mode: int, input_ids: torch.Tensor,
                            attention_mask: torch.Tensor, labels: Optional[torch.Tensor], speech_id: int, pad_id: int,
                            ignore_id: int = -100, pending: Optional[PendingSplicePlan] = None):
    """Step 4 on PACKED audio rows ``[sum M_b, H]`` (no [B, max M_b, H] padding in between); ``text_src`` is the
    embedding table (``text_mode=1``: embed_tokens lookup fused, ps-slm.py:525,654) or ``inputs_embeds``
    (``text_mode=0``).  Same 5-tuple and ValueErrors as ``_merge_input_ids_with_audio_features``
    (ps-slm.py:679-873); differentiable w.r.t. ``audio_rows``.  ``pending``: a plan begun before the audio rows
    were computed (``begin_splice_plan``)."""
    if audio_rows.dtype != text_src.dtype:
        audio_rows = audio_rows.to(text_src.dtype)
    if text_mode == 1 and torch.is_grad_enabled() and text_src.requires_grad:
        # embed_tokens is being trained (freeze_llm=False / PEFT "embs are hot", ps-slm.py:119-123): the lookup stays a
        # differentiable torch op (ps-slm.py:525) and the splice hands the gradient of its text rows back to it
        text_src, text_mode = torch.nn.functional.embedding(input_ids, text_src), 0
    if pending is None:
        pending = begin_splice_plan(input_ids, attention_mask, num_audio_tokens, speech_id)
    p, spliced_len = pending.finish()
    if torch.is_grad_enabled() and (audio_rows.requires_grad or text_src.requires_grad):
        from .autograd import SpliceFunction
        return SpliceFunction.apply(audio_rows, text_src, p, spliced_len, text_mode, 0,
                                    max_audio_tokens, labels, pad_id, ignore_id)
    return ops.splice_scatter(p, spliced_len, text_src, text_mode, audio_rows, 0, max_audio_tokens,
                              labels, pad_id, ignore_id, left_padding=p.left_padding)


def _raise_splice_errors(hdr, attention_mask, num_audios):
    if int(hdr[L.SH_ERR_BOTH_SIDES]):
        raise ValueError(f"both side of attention_mask has zero, invalid. {attention_mask}")
    if int(hdr[L.SH_N_SPEECH]) != num_audios and num_audios != 1:
        raise ValueError("shape mismatch: %d <speech> tokens for %d audios" % (int(hdr[L.SH_N_SPEECH]), num_audios))
    if int(hdr[L.SH_TOTAL_SLOTS]) != int(hdr[L.SH_TOTAL_AUDIO]):
        raise ValueError(
            f"The input provided to the model are wrong. The number of audio tokens is {int(hdr[L.SH_N_SPEECH])} while"
            f" the number of audio given to the model is {num_audios}. This prevents correct indexing and breaks batch generation.")


CACHE_GENERATION = [0]
VERIFY_CACHES = True      # validate cached weight copies against a content fingerprint of the live parameters


def invalidate_caches():
    """Drop every cached bf16 / folded weight copy (all ``ProjectorCache`` objects).  Called by ``Module.train()`` /
    ``.eval()`` of the bridge's modules; call it yourself after changing parameters through raw storage."""
    CACHE_GENERATION[0] += 1


class ProjectorCache:
    """bf16 / folded copies of (frozen or evaluation-time) weights.

    A copy is rebuilt when (a) a parameter's address, shape or torch version counter changes, (b) ``invalidate_caches()``
    ran since it was built — every ``train()`` / ``eval()`` of the bridge's modules does that — or (c) ``verify=True`` and
    the content fingerprint of the live parameters (tasu_fingerprint, one tiny kernel + one scalar read-back) differs
    from the one taken when the copy was built.  (c) is what catches optimizers that update parameters through flat
    buffers and ``.data.copy_`` (DeepSpeed ZeRO-1/2, finetune_deepspeed.py:147-149): neither address nor version
    counter moves there.  ``fresh=True`` (training forwards) never caches."""

    def __init__(self):
        self._key = None
        self.data = None
        self.fp = None
        self.builds = 0

    def get(self, params, builder, fresh=False, verify=False):
        if fresh:
            self._key, self.data, self.fp = None, None, None
            self.builds += 1
            return builder()
        live = [p for p in params if p is not None]
        key = (CACHE_GENERATION[0],) + tuple((p.data_ptr(), p._version, tuple(p.shape)) for p in live)
        fp = None
        if verify and VERIFY_CACHES and live and live[0].is_cuda:
            fp = int(ops.fingerprint(live[:8]).item())
        if key != self._key or (fp is not None and self.fp is not None and fp != self.fp):
            self.data = builder()
            self._key = key
            self.builds += 1
        if fp is not None:
            self.fp = fp
        return self.data


def cast_weight_bf16(w: torch.Tensor) -> torch.Tensor:
    """[N, K] weight → bf16 with a 64-element-padded pitch (zero padded)."""
    N, K = w.shape
    ld = ops.pad_to(K)
    if ld == K:
        out, _, _ = ops.cast_rows(w.detach(), torch.bfloat16)
        return out
    out = torch.zeros(N, ld, dtype=torch.bfloat16, device=w.device)
    tmp, _, _ = ops.cast_rows(w.detach(), torch.bfloat16)
    out[:, :K] = tmp
    return out


def _cap(n: int, q: int = 2048) -> int:
    """Allocation size for a data-dependent row count: rounded up so the caching allocator sees a handful of
    distinct sizes instead of a new one per batch (no cudaMalloc/cudaFree churn in steady state)."""
    return max(q, (n + q - 1) // q * q)


def linear_silu_forward(x_bf16: torch.Tensor, rows: int, K: int, mean: torch.Tensor, rstd: torch.Tensor,
                        w1g: torch.Tensor, colsum: torch.Tensor, dbias: torch.Tensor, w2: torch.Tensor,
                        b2: torch.Tensor, out_dtype: torch.dtype, simt: bool = False, stage=None,
                        m_dev: Optional[torch.Tensor] = None, streamk: bool = False) -> torch.Tensor:
    """LayerNorm → Linear → SiLU → Linear of projector.py:149-151 as two tensor-core GEMMs:
    GEMM-1 runs on the raw rows with the LayerNorm folded into its epilogue.
    ``streamk``: GEMM-1 through ``tasu_gemm_bf16_tn_streamk`` — the ragged last wave of tiles cut along K (the row count
    is data dependent; 3.57 waves at the headline batch: −4.4 %, profiles/r02n_gemm_ab.md)."""
    Hb, H = w1g.shape[0], w2.shape[0]
    dev = x_bf16.device
    import contextlib
    stage = stage or (lambda name: contextlib.nullcontext())
    h1 = torch.empty(_cap(rows), Hb, dtype=torch.bfloat16, device=dev)[:rows]
    with stage("projector_gemm1"):
        if streamk and not simt:
            ops.gemm_bf16_tn_streamk(x_bf16, w1g, rows, Hb, K, h1, L.EPI_LNFOLD_SILU, dbias, rstd, mean, colsum, m_dev=m_dev)
        else:
            ops.gemm_bf16_tn(x_bf16, w1g, rows, Hb, K, h1, L.EPI_LNFOLD_SILU, dbias, rstd, mean, colsum, simt=simt,
                             m_dev=m_dev)
    y = torch.empty(_cap(rows), H, dtype=out_dtype, device=dev)[:rows]
    with stage("projector_gemm2"):
        ops.gemm_bf16_tn(h1, w2, rows, H, Hb, y, L.EPI_BIAS, b2, simt=simt, m_dev=m_dev)
    return y


class _Stage:
    """Context manager recording a CUDA-event pair on the current stream around one stage."""

    def __init__(self, owner, name):
        self.owner, self.name = owner, name

    def __enter__(self):
        if self.owner.profile:
            self.t0 = torch.cuda.Event(enable_timing=True)
            self.t1 = torch.cuda.Event(enable_timing=True)
            self.t0.record()
        return self

    def __exit__(self, *exc):
        if self.owner.profile:
            self.t1.record()
            self.owner.events.append((self.name, self.t0, self.t1))
        return False


class TasuBridge:
    """Fused inference bridge: encoder output → inputs_embeds.

    ``raw_encoder_out [B, T+4, 512]`` → ctc_lo GEMM (tcgen05) → per-frame softmax stats /
    argmax → collapse plan → splice plan → ONE header read-back → softmax + segmented
    mean-pool of the kept frames (bf16, LayerNorm stats fused) → LN-folded GEMM-1 + SiLU →
    GEMM-2 → gather/scatter splice with the embed_tokens lookup fused.
    Restates the dispatch of ps-slm.py:581-658 for ctc_posterior=True, voca_trans=False,
    gt_emb=False, do_psd=True, encoder_projector='linear-silu'."""

    N_PREFIX = 4     # language / event / emotion / textnorm query frames (ps-slm.py:442-443,452-454)

    def __init__(self, w_ctc: torch.Tensor, b_ctc: Optional[torch.Tensor], projector, embed_table: torch.Tensor,
                 speech_id: int, pad_id: int, ignore_id: int = -100, blank_id: int = 0,
                 blank_threshold: float = BLANK_THRESHOLD, ln_eps: float = 1e-5):
        self.w_ctc, self.b_ctc = w_ctc, b_ctc
        self.projector = projector
        self.embed_table = embed_table
        self.speech_id, self.pad_id, self.ignore_id = int(speech_id), int(pad_id), int(ignore_id)
        self.blank_id, self.blank_threshold, self.ln_eps = int(blank_id), float(blank_threshold), float(ln_eps)
        self._ctc_cache = ProjectorCache()
        self._fp = None               # content fingerprint of the parameters the cached weight copies were made from
        self._capacity = {}           # (B, T) → (kept-frame rows, packed rows) the tail buffers are sized for
        self.last_counts = {}
        # Exact decisions (default): the frames whose greedy decisions — argmax (ps-slm.py:265), strict fp32 blank
        # threshold (:295-297) — lie inside the rounding-error bound of the bf16 head are recomputed with fp32 FMAs from the
        # fp32 weights before the collapse plan (csrc/refine.cu), so every index / run boundary / kept row equals the fp32
        # reference.  Per-frame bound, uncapped list, cost proportional to the number of such frames (3 small launches when
        # there are none).  False: decisions straight from the bf16 head.
        self.exact_decisions = True
        self.last_ambiguous = None        # device int32[1]: frames refined by the last call (exact_decisions)
        self.last_multi = None            # device int32[1]: multi-frame candidates pooled by the last call
        self._ctc_exact_cache = ProjectorCache()
        self.materialize_logits = False   # True: ctc_lo writes fp32 logits to HBM + streaming stats kernel (round-1a path)
        # "bf16" (default): bf16 operands, fp32 accumulation — posteriors / embeddings within 1e-2 of the fp32 reference.
        # "fp32x3": reference numerics (fp32, conf/ds_config.json:12-14): the kept frames' logits and both projector
        # contractions run as three-term bf16 splits on the tensor cores, softmax / pooling / LayerNorm in fp32 —
        # embeddings within 1e-5 of the fp32 reference, ~6 x the tensor work on the kept frames (two-phase: header first)
        self.precision = "bf16"
        self._ctc_split_cache = ProjectorCache()
        # projector GEMM-1 with the stream-K tail (default; TASU_GEMM_STREAMK=0 selects the plain persistent kernel)
        self.streamk_gemm1 = os.environ.get("TASU_GEMM_STREAMK", "1") != "0"
        # Grouped kept-frame layout (default; TASU_GROUPED_POOL=0 selects the plain layout + tail pooling): runs of 2-4
        # frames are averaged inside the epilogue of the kept-frame GEMM (csrc/grouped.cu)
        self.grouped_pool = os.environ.get("TASU_GROUPED_POOL", "1") != "0"
        self.profile = False          # when True, CUDA events bracket every stage (bench roofline)
        self.events = []              # [(stage name, start event, end event)] of the profiled calls

    def _stage(self, name):
        return _Stage(self, name)

    def _pool_kept(self, x2, st, plan, B, T, Denc, V, cap_f, cap_o, w_ctc, b_ctc):
        """Pass 2 on capacity-sized buffers with device-side row counts (no host sync inside): gather kept rows →
        softmax-epilogue CTC GEMM over the kept frames → in-place tail pooling.  Row r of the compact matrix is the first
        frame of packed candidate r, so single-frame candidates (the majority) are final as the GEMM writes them; only
        multi-frame runs are averaged afterwards.  → (pooled bf16 [cap_f, pad64(V)], LayerNorm mean, rstd [cap_o])."""
        dev = x2.device
        ldk = ops.pad_to(V)
        with self._stage("gather_kept_rows"):
            xg, g_max, g_inv, pk_len, tail_src, multi, mean, rstd = ops.gather_kept_rows(
                x2, B, T, self.N_PREFIX, Denc, V, plan, st, cap_f, cap_o, self.ln_eps)
        self.last_multi = multi[0:1]      # device int32[1]: multi-frame candidates of the last call (pool_tail's work list)
        pooled = torch.empty(cap_f, ldk, dtype=torch.bfloat16, device=dev)
        with self._stage("ctc_softmax_gemm"):
            ops.gemm_bf16_tn(xg, w_ctc, cap_f, V, Denc, pooled, L.EPI_SOFTMAX, b_ctc, g_inv, g_max,
                             m_dev=plan.counts[2:3])
        with self._stage("pool_tail"):
            ops.pool_tail(pooled, V, cap_o, pk_len, tail_src, multi, mean, rstd, self.ln_eps)
        return pooled, mean, rstd

    def _pool_kept_grouped(self, x2, st, plan, B, T, Denc, V, cap_f, cap_o, w_ctc, b_ctc, max_proj):
        """Pass 2 on the grouped layout (include/tasu_bridge.h, step 2b'): the frames of a run of 2-4 frames sit on adjacent
        rows of the kept-frame GEMM and are averaged in its epilogue — their per-frame probabilities are never written,
        ``pool_tail`` only sees runs of more than 4 frames.  Pooled rows come out in class order (``g.perm``: packed
        candidate → pooled row).  → (pooled bf16 [cap_a, pad64(V)], GroupedRows)."""
        with self._stage("gather_kept_rows"):
            g = ops.gather_kept_rows_grouped(x2, B, T, self.N_PREFIX, Denc, V, plan, st, cap_f, cap_o, max_proj, self.ln_eps)
        self.last_multi = g.multi[0:1]    # device int32[1]: runs of more than 4 frames (pool_tail's work list)
        pooled = torch.empty(g.cap_p, ops.pad_to(V), dtype=torch.bfloat16, device=x2.device)
        with self._stage("ctc_softmax_gemm"):
            ops.gemm_softmax_grouped(g, w_ctc, b_ctc, V, Denc, pooled, self.ln_eps)
        with self._stage("pool_tail"):
            ops.pool_tail(pooled, V, min(cap_o, 1024), g.pk_len, g.tail_src, g.multi, g.mean, g.rstd, self.ln_eps)
        return pooled, g

    def _tail(self, x2, st, plan, B, T, Denc, V, cap_f, cap_o, w_ctc, b_ctc, w1g, colsum, dbias, w2, b2, out_dtype,
              keep_perm=False):
        """Pass 2 + projector (GEMM-1 with the LayerNorm folded in, GEMM-2) on capacity-sized buffers → audio rows in
        packed candidate order; with ``keep_perm`` (grouped layout) → ``(rows in class order, perm)`` for a consumer that
        reads through the permutation itself (the splice)."""
        if self.grouped_pool:
            rows = cap_o + 256                                          # pooled rows incl. the holes between the classes
            pooled, g = self._pool_kept_grouped(x2, st, plan, B, T, Denc, V, cap_f, cap_o, w_ctc, b_ctc, rows)
            y = linear_silu_forward(pooled, rows, V, g.mean, g.rstd, w1g, colsum, dbias, w2, b2, out_dtype,
                                    stage=self._stage, m_dev=g.lay[L.GL_OX:L.GL_OX + 1], streamk=self.streamk_gemm1)
            if keep_perm:
                return y, g.perm[:cap_o]
            with self._stage("unpermute_rows"):
                return ops.gather_rows(y, g.perm[:cap_o])               # back to packed candidate order
        if keep_perm:
            return self._tail(x2, st, plan, B, T, Denc, V, cap_f, cap_o, w_ctc, b_ctc, w1g, colsum, dbias, w2, b2,
                              out_dtype), None
        pooled, mean, rstd = self._pool_kept(x2, st, plan, B, T, Denc, V, cap_f, cap_o, w_ctc, b_ctc)
        return linear_silu_forward(pooled, cap_o, V, mean, rstd, w1g, colsum, dbias, w2, b2, out_dtype,
                                   stage=self._stage, m_dev=plan.counts[0:1], streamk=self.streamk_gemm1)

    def _tail_fp32(self, raw_encoder_out, plan, B, T, Denc, V, n_frames, n_out, max_len, b_ctc):
        """fp32-accurate pass 2 + projector (precision = "fp32x3"), sized exactly: gather the kept frames' fp32 encoder
        rows in natural order → logits with the three-term bf16 split GEMM (~1e-6 of fp32) → fp32 row statistics →
        softmax fused into the segmented mean-pool (fp32 rows + LayerNorm statistics) → fp32-accurate projector."""
        dev = raw_encoder_out.device
        H = self.embed_table.shape[1]
        if n_out == 0:
            return torch.zeros(0, H, dtype=torch.float32, device=dev)
        rows = raw_encoder_out.reshape(B * (T + self.N_PREFIX), Denc)
        rows = rows if rows.dtype == torch.float32 else rows.float()
        with self._stage("fp32_gather_kept"):
            frame_row, seg_src = ops.kept_frame_index(plan, self.N_PREFIX, n_frames, n_out)
            xg = ops.gather_rows(rows, frame_row)
        with self._stage("fp32_ctc_logits"):
            w_split, k_split = self._ctc_split_cache.get(
                [self.w_ctc], lambda: ops.split_bf16x3(self.w_ctc.detach().float(), 1)[:2], verify=True)
            xs, kx, _, _, _ = ops.split_bf16x3(xg, 0)
            ldv = ops.pad_to(V, 4)
            logits = torch.empty(n_frames, ldv, dtype=torch.float32, device=dev)
            ops.gemm_fp32x3(xs, w_split, n_frames, V, kx, logits, L.EPI_BIAS, b_ctc)
        with self._stage("fp32_softmax_meanpool"):
            st2 = ops.frame_stats(logits.view(1, n_frames, ldv)[:, :, :V], L.INPUT_LOGITS, self.blank_id)
            pooled = torch.empty(n_out, ldv, dtype=torch.float32, device=dev)
            ops.segment_meanpool(logits, plan, 0, max_len, n_out, pooled, ldv, softmax=st2, seg_src=seg_src, feat_dim=V)
        with self._stage("fp32_projector"):
            return self.projector.forward_rows_fp32x3(pooled[:, :V])

    def _speculative_capacity(self, B, T):
        """Row capacities for a tail enqueued BEFORE the host knows the batch's counts: 1.25 x the high-water marks of
        earlier batches of this shape, or None for the first batch of a shape — that one call reads the header first and
        sizes its buffers exactly (a worst-case B*T-row probability matrix would be 1.6 GB at 64 x 30 s and 51 GB at
        1024 x 60 s)."""
        hw_f, hw_o = self._capacity.get((B, T), (0, 0))
        if not hw_f:
            return None
        return _cap(int(1.25 * hw_f)), _cap(int(1.25 * hw_o))

    def _header_slot(self):
        """Ring of pinned host header buffers (collapse + splice words), one per in-flight call."""
        if not hasattr(self, "_hdr_ring"):
            self._hdr_ring = [torch.zeros(L.CH_WORDS + L.SH_WORDS + 1, dtype=torch.int64).pin_memory() for _ in range(4)]
            self._hdr_i = 0
        self._hdr_i = (self._hdr_i + 1) % len(self._hdr_ring)
        return self._hdr_ring[self._hdr_i]

    def _ctc_weights(self):
        def build():
            w = cast_weight_bf16(self.w_ctc)
            b = (self.b_ctc.detach().float().contiguous() if self.b_ctc is not None
                 else torch.zeros(self.w_ctc.shape[0], dtype=torch.float32, device=self.w_ctc.device))
            return w, b
        return self._ctc_cache.get([self.w_ctc, self.b_ctc], build)

    def _ctc_exact_weights(self):
        """fp32 weights of the CTC head + max_v ||w_v|| (device scalar) for the exact-decision refinement."""
        def build():
            w = self.w_ctc.detach()
            w = w if (w.dtype == torch.float32 and w.is_contiguous()) else w.float().contiguous()
            return w, ops.row_norm_max(w)
        return self._ctc_exact_cache.get([self.w_ctc], build)

    def _encoder_rows_bf16(self, raw_encoder_out):
        """[B, T+4, D] encoder output → bf16 rows [B*(T+4), pad64(D)] for the tensor cores (a no-op for a bf16 hand-over);
        with exact decisions the fp32 cast also yields the squared row norms of the error bound (``self._x_sumsq``)."""
        B, T4, Denc = raw_encoder_out.shape
        x2 = raw_encoder_out.reshape(B * T4, Denc)
        self._x_sumsq = None
        if x2.dtype == torch.bfloat16:
            return x2
        with self._stage("cast_encoder_out"):
            if self.exact_decisions and x2.dtype == torch.float32:
                x2, self._x_sumsq = ops.cast_rows_sumsq(x2, ops.pad_to(Denc))
            else:
                x2, _, _ = ops.cast_rows(x2, torch.bfloat16, ops.pad_to(Denc))
        return x2

    def _head_stats(self, raw_encoder_out, x2, lens, w_ctc, b_ctc, B, T, Denc, V):
        """(a1 + the decisions of a2) fused CTC head statistics, refined to fp32-exact decisions when asked for."""
        with self._stage("ctc_head_stats"):
            st = ops.ctc_head_stats(x2, w_ctc, b_ctc, B, T, self.N_PREFIX, V, Denc, self.blank_id)
        if self.exact_decisions:
            with self._stage("refine_ambiguous"):
                w32, wnorm = self._ctc_exact_weights()
                rows = raw_encoder_out.reshape(B * (T + self.N_PREFIX), Denc)
                if rows.dtype not in (torch.float32, torch.bfloat16):
                    rows = rows.float()
                self.last_ambiguous = ops.refine_ambiguous_frames(st, lens, rows, w32, b_ctc, wnorm, T, self.N_PREFIX, V,
                                                                  self.blank_id, self.blank_threshold,
                                                                  x_sumsq=getattr(self, "_x_sumsq", None))
        return st

    def _weight_params(self):
        fast = getattr(self.projector, "weight_params", None)  # direct attribute access: walking the module tree costs ~10 us
        return [self.w_ctc, self.b_ctc] + (fast() if fast is not None else [p for p in self.projector.parameters()][:6])

    def _cache_builds(self):
        return self._ctc_cache.builds + self._ctc_exact_cache.builds + self.projector._cache.builds

    def _check_fingerprint(self, fp: int, builds_before: int) -> bool:
        """True when the cached weight copies this call used are current.  The fingerprint of the live parameters was
        taken by the first kernel of the call and arrives with the plan header (no extra synchronisation)."""
        if not VERIFY_CACHES:
            return True
        rebuilt = self._cache_builds() != builds_before
        if self._fp is None or fp == self._fp or rebuilt:
            self._fp = fp
            return True
        # parameters changed behind torch's version counter (flat-buffer optimizer): drop the copies, redo the call
        self._fp = None
        invalidate_caches()
        return False

    @torch.no_grad()
    def __call__(self, raw_encoder_out: torch.Tensor, raw_encoder_out_lens: torch.Tensor,
                 input_ids: torch.Tensor, attention_mask: torch.Tensor, labels: Optional[torch.Tensor] = None,
                 want_ids: bool = False):
        for _ in range(2):
            out = self._run(raw_encoder_out, raw_encoder_out_lens, input_ids, attention_mask, labels, want_ids)
            if out is not None:
                return out
        raise L.TasuError("projector / CTC-head parameters keep changing while the bridge runs")

    def _run(self, raw_encoder_out, raw_encoder_out_lens, input_ids, attention_mask, labels, want_ids):
        B, T4, Denc = raw_encoder_out.shape
        T = T4 - self.N_PREFIX
        V = self.w_ctc.shape[0]
        dev = raw_encoder_out.device
        builds_before = self._cache_builds()
        header = self._header_slot()
        if VERIFY_CACHES:
            ops.fingerprint(self._weight_params(), out=header[L.CH_WORDS + L.SH_WORDS:])
        w_ctc, b_ctc = self._ctc_weights()
        w1g, colsum, dbias, w2, b2 = self.projector.folded_weights()
        out_dtype = self.embed_table.dtype

        # splice row statistics only depend on the prompt: issue them first
        with self._stage("splice_plan"):
            sp = ops.splice_rowstat(input_ids, attention_mask, self.speech_id)

        x2 = self._encoder_rows_bf16(raw_encoder_out)
        lens = torch.clamp(raw_encoder_out_lens.to(device=dev, dtype=torch.int64) - self.N_PREFIX, min=0)
        # The plan headers are written by the kernels straight into pinned (UVA-mapped) host memory: the one
        # device→host hand-off of the step needs no copy-engine transfer, so it cannot queue behind the bulk
        # D2H of the previous batch when calls are pipelined (HostPipeline).
        ldk = ops.pad_to(V)

        if self.materialize_logits:
            # (a1) ctc_lo on the tensor cores: logits [B*(T+4), ldv] fp32, then one streaming stats pass
            ldv = ops.pad_to(V, 4)
            logits = torch.empty(B * T4, ldv, dtype=torch.float32, device=dev)
            with self._stage("ctc_lo_gemm"):
                ops.gemm_bf16_tn(x2, w_ctc, B * T4, V, Denc, logits, L.EPI_BIAS, b_ctc)
            post_view = logits.view(B, T4, ldv)[:, self.N_PREFIX:, :V]      # ps-slm.py:583 (logits, not probs)
            with self._stage("frame_stats"):
                st = ops.frame_stats(post_view, L.INPUT_LOGITS, self.blank_id, lens)
        else:
            # (a1+a2) fused: logits live only in TMEM, softmax statistics come out of the GEMM epilogue
            st = self._head_stats(raw_encoder_out, x2, lens, w_ctc, b_ctc, B, T, Denc, V)

        # (a2) collapse plan, (a8) splice plan; one header for both
        with self._stage("collapse_plan"):
            plan = ops.collapse_plan(st, lens, self.blank_id, self.blank_threshold, header=header[:L.CH_WORDS])
        with self._stage("splice_plan"):
            ops.splice_plan(sp, plan.new_lens, self.projector.k, header=header[L.CH_WORDS:])
        ev = torch.cuda.Event()
        ev.record()                                                     # header is complete when this event fires
        audio_cap, audio_perm, cap_f, cap_o = None, None, 0, 0
        fp32 = self.precision == "fp32x3" and not self.materialize_logits
        caps = None if (self.materialize_logits or fp32) else self._speculative_capacity(B, T)
        if caps is not None:
            # The whole tail is enqueued BEFORE the host looks at the header: buffers are sized by capacity
            # (high-water mark of earlier calls) and every kernel takes its live row count from device memory, so the
            # GPU never idles waiting for the host.
            cap_f, cap_o = caps
            audio_cap, audio_perm = self._tail(x2, st, plan, B, T, Denc, V, cap_f, cap_o, w_ctc, b_ctc, w1g, colsum, dbias,
                                               w2, b2, out_dtype, keep_perm=True)
        ev.synchronize()                                                # the single device→host hand-off
        hdr = header.clone()
        if not self._check_fingerprint(int(hdr[L.CH_WORDS + L.SH_WORDS]), builds_before):
            return None                                                 # stale weight copies: the caller redoes the call
        n_out, max_len = int(hdr[L.CH_N_OUT]), int(hdr[L.CH_MAX_LEN])
        n_frames = int(hdr[L.CH_KEPT_FRAMES])
        shdr = hdr[L.CH_WORDS:L.CH_WORDS + L.SH_WORDS]
        _raise_splice_errors(shdr, attention_mask, B)
        spliced_len = int(shdr[L.SH_SPLICED_LEN])

        if self.materialize_logits:
            if n_out > 0:
                pooled = torch.empty(_cap(n_out), ldk, dtype=torch.bfloat16, device=dev)[:n_out]
                mean = torch.empty(n_out, dtype=torch.float32, device=dev)
                rstd = torch.empty(n_out, dtype=torch.float32, device=dev)
                with self._stage("softmax_meanpool"):
                    ops.segment_meanpool(post_view, plan, 0, max_len, n_out, pooled, ldk, softmax=st,
                                         ln_mean=mean, ln_rstd=rstd, ln_eps=self.ln_eps)
                audio = linear_silu_forward(pooled, n_out, V, mean, rstd, w1g, colsum, dbias, w2, b2, out_dtype,
                                            stage=self._stage)
            else:
                audio = torch.empty(0, self.embed_table.shape[1], dtype=out_dtype, device=dev)
        elif fp32:
            audio = self._tail_fp32(raw_encoder_out, plan, B, T, Denc, V, n_frames, n_out, max_len, b_ctc).to(out_dtype)
        else:
            if audio_cap is None or n_frames > cap_f or n_out > cap_o:
                # first batch of this shape, or capacity exceeded (rare): the tail runs now, sized exactly
                cap_f, cap_o = _cap(n_frames), _cap(n_out)
                audio_cap, audio_perm = self._tail(x2, st, plan, B, T, Denc, V, cap_f, cap_o, w_ctc, b_ctc, w1g, colsum,
                                                   dbias, w2, b2, out_dtype, keep_perm=True)
            hw_f, hw_o = self._capacity.get((B, T), (0, 0))
            self._capacity[(B, T)] = (max(hw_f, n_frames), max(hw_o, n_out))
            # grouped layout: the rows stay in class order, the splice reads them through the permutation
            audio = audio_cap if audio_perm is not None else audio_cap[:n_out]
        # (a7+a8) splice with the embedding lookup fused
        with self._stage("splice_scatter"):
            emb, mask, out_labels, pos, fids = ops.splice_scatter(
                sp, spliced_len, self.embed_table, 1, audio, 0, max_len, labels, self.pad_id, self.ignore_id,
                want_ids=want_ids, left_padding=int(shdr[L.SH_LEFT_PADDING]), audio_perm=audio_perm)
        self.last_counts = {"n_in": int(B * T), "n_out": n_out, "max_len": max_len, "spliced_len": spliced_len,
                            "kept_frames": int(hdr[L.CH_KEPT_FRAMES])}
        return emb, mask, out_labels, pos, plan.new_lens

    # ------------------------------------------------------------------ two-phase API (cross-rank packing)
    @torch.no_grad()
    def compress_project_async(self, raw_encoder_out: torch.Tensor, raw_encoder_out_lens: torch.Tensor) -> "PendingCompress":
        """Steps 1b-3, enqueued without a host synchronisation: head statistics → (exact decisions) → collapse plan →
        speculative tail on capacity-sized buffers (as in ``__call__``).  ``finish()`` waits for the header only."""
        B, T4, Denc = raw_encoder_out.shape
        T = T4 - self.N_PREFIX
        V = self.w_ctc.shape[0]
        dev = raw_encoder_out.device
        pend = PendingCompress()
        pend.bridge, pend.inputs = self, (raw_encoder_out, raw_encoder_out_lens)
        pend.builds_before = self._cache_builds()
        header = self._header_slot()
        if VERIFY_CACHES:
            ops.fingerprint(self._weight_params(), out=header[L.CH_WORDS + L.SH_WORDS:])
        w_ctc, b_ctc = self._ctc_weights()
        proj_w = self.projector.folded_weights()
        out_dtype = self.embed_table.dtype
        x2 = self._encoder_rows_bf16(raw_encoder_out)
        lens = torch.clamp(raw_encoder_out_lens.to(device=dev, dtype=torch.int64) - self.N_PREFIX, min=0)
        st = self._head_stats(raw_encoder_out, x2, lens, w_ctc, b_ctc, B, T, Denc, V)
        plan = ops.collapse_plan(st, lens, self.blank_id, self.blank_threshold, header=header[:L.CH_WORDS])
        ev = torch.cuda.Event()
        ev.record()
        pend.tail_args = (x2, st, plan, B, T, Denc, V)
        pend.weights = (w_ctc, b_ctc) + tuple(proj_w) + (out_dtype,)
        caps = self._speculative_capacity(B, T)
        pend.audio_cap, pend.cap_f, pend.cap_o = None, 0, 0
        if caps is not None:
            pend.cap_f, pend.cap_o = caps
            pend.audio_cap = self._tail(x2, st, plan, B, T, Denc, V, caps[0], caps[1], w_ctc, b_ctc, *proj_w, out_dtype)
        pend.header, pend.event, pend.plan = header, ev, plan
        return pend

    @torch.no_grad()
    def compress_pooled(self, raw_encoder_out: torch.Tensor, raw_encoder_out_lens: torch.Tensor):
        """Steps 1b-2 only, for TRAINING on audio batches (ps-slm.py:450-454, :469-473): the fused head statistics,
        exact decisions, collapse plan, kept-frame softmax GEMM and tail pooling — the no-grad part of the path, the
        ``[B, T, 25055]`` posterior never exists — and the pooled posterior rows are handed to the differentiable
        projector (``autograd.linear_silu_train_rows``).  Reads the plan header first, so every buffer has its exact size.
        → (pooled bf16 ``[sum M_b, pad64(V)]``, LayerNorm mean, rstd ``[sum M_b]``, ``new_lens [B]`` int64, ``max_b M_b``)."""
        B, T4, Denc = raw_encoder_out.shape
        T = T4 - self.N_PREFIX
        V = self.w_ctc.shape[0]
        dev = raw_encoder_out.device
        for _ in range(2):
            builds_before = self._cache_builds()
            header = self._header_slot()
            if VERIFY_CACHES:
                ops.fingerprint(self._weight_params(), out=header[L.CH_WORDS + L.SH_WORDS:])
            w_ctc, b_ctc = self._ctc_weights()
            x2 = self._encoder_rows_bf16(raw_encoder_out)
            lens = torch.clamp(raw_encoder_out_lens.to(device=dev, dtype=torch.int64) - self.N_PREFIX, min=0)
            st = self._head_stats(raw_encoder_out, x2, lens, w_ctc, b_ctc, B, T, Denc, V)
            plan = ops.collapse_plan(st, lens, self.blank_id, self.blank_threshold, header=header[:L.CH_WORDS])
            ev = torch.cuda.Event()
            ev.record()
            ev.synchronize()
            hdr = header.clone()
            if self._check_fingerprint(int(hdr[L.CH_WORDS + L.SH_WORDS]), builds_before):
                break
        n_out, max_len, n_frames = int(hdr[L.CH_N_OUT]), int(hdr[L.CH_MAX_LEN]), int(hdr[L.CH_KEPT_FRAMES])
        if n_out == 0:
            z = torch.zeros(0, dtype=torch.float32, device=dev)
            return torch.zeros(0, ops.pad_to(V), dtype=torch.bfloat16, device=dev), z, z, plan.new_lens, 0
        pooled, mean, rstd = self._pool_kept(x2, st, plan, B, T, Denc, V, _cap(n_frames), _cap(n_out), w_ctc, b_ctc)
        self.last_counts = {"n_in": int(B * T), "n_out": n_out, "max_len": max_len, "kept_frames": n_frames}
        return pooled[:n_out], mean[:n_out], rstd[:n_out], plan.new_lens, max_len

    def compress_project(self, raw_encoder_out: torch.Tensor, raw_encoder_out_lens: torch.Tensor):
        """Steps 1b-3 only: → (audio rows packed ``[sum M_b, H]``, ``new_lens [B]`` int64, ``max_b M_b``).
        Used when compressed sequences are exchanged between ranks before the splice (dist.gather_packed)."""
        return self.compress_project_async(raw_encoder_out, raw_encoder_out_lens).finish()

    @torch.no_grad()
    def splice(self, audio_rows: torch.Tensor, new_lens: torch.Tensor, input_ids: torch.Tensor,
               attention_mask: torch.Tensor, labels: Optional[torch.Tensor] = None, want_ids: bool = False):
        """Step 4 on packed audio rows (any origin: this rank's or gathered from all ranks)."""
        sp = ops.splice_rowstat(input_ids, attention_mask, self.speech_id)
        header = self._header_slot()
        ops.splice_plan(sp, new_lens, self.projector.k, header=header[L.CH_WORDS:])
        ev = torch.cuda.Event()
        ev.record()
        ev.synchronize()
        shdr = header[L.CH_WORDS:].clone()
        _raise_splice_errors(shdr, attention_mask, new_lens.numel())
        return ops.splice_scatter(sp, int(shdr[L.SH_SPLICED_LEN]), self.embed_table, 1, audio_rows, 0, 0, labels,
                                  self.pad_id, self.ignore_id, want_ids=want_ids,
                                  left_padding=int(shdr[L.SH_LEFT_PADDING]))


class PendingCompress:
    """Handle of ``TasuBridge.compress_project_async``: the kernels are enqueued, ``finish()`` waits for the 96-byte plan
    header, validates the cached weight copies, redoes the tail if a capacity was exceeded and returns
    ``(audio rows [sum M_b, H] — a view of a capacity-sized buffer —, new_lens [B] int64, max_b M_b)``."""
    __slots__ = ("bridge", "inputs", "builds_before", "tail_args", "weights", "audio_cap", "cap_f", "cap_o", "header",
                 "event", "plan")

    def finish(self):
        br = self.bridge
        self.event.synchronize()
        hdr = self.header.clone()
        if not br._check_fingerprint(int(hdr[L.CH_WORDS + L.SH_WORDS]), self.builds_before):
            return br.compress_project(*self.inputs)                    # stale weight copies: redo with fresh ones
        n_out, max_len, n_frames = int(hdr[L.CH_N_OUT]), int(hdr[L.CH_MAX_LEN]), int(hdr[L.CH_KEPT_FRAMES])
        x2, st, plan, B, T, Denc, V = self.tail_args
        if self.audio_cap is None or n_frames > self.cap_f or n_out > self.cap_o:
            # first batch of this shape, or capacity exceeded (rare): the tail runs now, sized exactly
            self.audio_cap = br._tail(x2, st, plan, B, T, Denc, V, _cap(n_frames), _cap(n_out), *self.weights)
        hw_f, hw_o = br._capacity.get((B, T), (0, 0))
        br._capacity[(B, T)] = (max(hw_f, n_frames), max(hw_o, n_out))
        br.last_counts = {"n_in": int(B * T), "n_out": n_out, "max_len": max_len, "kept_frames": n_frames}
        return self.audio_cap[:n_out], self.plan.new_lens, max_len


class HostPipeline:
    """End-to-end entry for HOST buffers (the call a serving loop makes): pinned host batches in,
    pinned host results out, with the H2D copy of batch i+1 and the D2H copy of batch i-1 overlapped
    with the kernels of batch i on three CUDA streams.

        pipe = HostPipeline(bridge)
        for emb, mask, pos, new_lens in pipe.run(batches):   # batches: (raw, raw_lens, input_ids, attention_mask)
            ...                                              # pinned CPU tensors, valid until the next iteration
    """

    def __init__(self, bridge: "TasuBridge", device=None, compute_streams: int = 1):
        self.bridge = bridge
        self.device = device if device is not None else bridge.embed_table.device
        self.s_in = torch.cuda.Stream(self.device)
        # batch i runs on compute stream i % N: with N = 2 the small HBM- / latency-bound kernels of one batch run in the
        # tails of the other batch's persistent GEMMs (as in bench.py's device-resident loop)
        self.s_comps = [torch.cuda.Stream(self.device) for _ in range(max(1, int(compute_streams)))]
        self.s_comp = self.s_comps[0]
        self.s_out = torch.cuda.Stream(self.device)
        self._host_out = {}
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _upload(self, batch):
        with torch.cuda.stream(self.s_in):
            dev = []
            for t in batch:
                if not t.is_pinned():
                    t = t.pin_memory()
                d = t.to(self.device, non_blocking=True)
                for sc in self.s_comps:
                    d.record_stream(sc)
                dev.append(d)
            ev = torch.cuda.Event()
            ev.record(self.s_in)
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in batch)
        return dev, ev

    def _download(self, outs, slot, s_comp):
        key = (slot,) + tuple((tuple(o.shape), o.dtype) for o in outs)
        if key not in self._host_out:
            self._host_out[key] = tuple(torch.empty(o.shape, dtype=o.dtype, pin_memory=True) for o in outs)
        host = self._host_out[key]
        ev_c = torch.cuda.Event()
        ev_c.record(s_comp)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(ev_c)
            for o, h in zip(outs, host):
                o.record_stream(self.s_out)
                h.copy_(o, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.s_out)
        self.d2h_bytes = sum(o.numel() * o.element_size() for o in outs)
        return host, ev

    def run(self, batches):
        it = iter(batches)
        try:
            nxt = self._upload(next(it))
        except StopIteration:
            return
        pending = None
        i = 0
        while nxt is not None:
            dev, ev_in = nxt
            try:
                nxt = self._upload(next(it))          # H2D of the next batch runs under this batch's kernels
            except StopIteration:
                nxt = None
            s_comp = self.s_comps[i % len(self.s_comps)]
            with torch.cuda.stream(s_comp):
                s_comp.wait_event(ev_in)
                emb, mask, _, pos, new_lens = self.bridge(*dev)
            done = self._download((emb, mask, pos, new_lens), i & 1, s_comp)
            if pending is not None:
                pending[1].synchronize()
                yield pending[0]
            pending = done
            i += 1
        if pending is not None:
            pending[1].synchronize()
            yield pending[0]
