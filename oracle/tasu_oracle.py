"""Torch-CPU restatement of the TASU bridge hot path (the parity oracle).

TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.  Never imported by the
product package.  Every function cites the reference lines it restates
(paths relative to /root/reference/).  Pinned against the real reference by
tests/test_oracle_vs_reference.py (container) and tests/golden/*.npz (anywhere).

Conventions: B utterances, T frames, V CTC vocab, D feature width, L_b valid
frames, M_b compressed length, S prompt width, S' spliced width, H LLM width.
"""
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BLANK_THRESHOLD = 0.90


# --------------------------------------------------------------------------
# (a1) CTC head + posterior — Multitask/model/ps-slm.py:450-454 / :581-585
# --------------------------------------------------------------------------
def ctc_head_posterior(raw_encoder_out: torch.Tensor, raw_lens: torch.Tensor,
                       w_ctc: torch.Tensor, b_ctc: Optional[torch.Tensor]):
    """softmax(ctc_lo(x))[:, 4:], lens = clamp(raw_lens - 4, 0).

    ``ctc_lo`` is funasr's CTC linear (un-vendored, unpinned dependency; call
    sites ps-slm.py:450,581, SenseVoice.py:619) restated as F.linear —
    parity unpinned at this boundary."""
    logits = F.linear(raw_encoder_out, w_ctc, b_ctc)
    post = torch.softmax(logits, dim=-1)[:, 4:, :]
    lens = torch.clamp(raw_lens - 4, min=0)
    return post, lens


# --------------------------------------------------------------------------
# (a2) PSD — Multitask/model/ps-slm.py:237-317
# --------------------------------------------------------------------------
def psd_plan(ctc_posterior: torch.Tensor, lens: Sequence[int], blank_id: int = 0,
             blank_threshold: float = BLANK_THRESHOLD):
    """Integer plan of PSD: for every utterance the list of kept candidates
    ``(start, length)`` plus per-frame greedy ids and per-candidate fp32 scores.

    ps-slm.py:256-257 (log-prob detection over the WHOLE tensor), :265 argmax
    (first max index), :270-288 runs (blank frames stand alone, non-blank runs
    merge), :291 scores collected to fp32, :295 strict ``<`` threshold."""
    B, T, V = ctc_posterior.shape
    is_log = bool(ctc_posterior.max() <= 0) if ctc_posterior.numel() else False
    ids_all = np.zeros((B, T), dtype=np.int64)
    out = []
    for b in range(B):
        L = int(lens[b])
        segs, scores = [], []
        if L > 0:
            rows = ctc_posterior[b, :L]
            ids = rows.argmax(dim=-1).numpy()
            ids_all[b, :L] = ids
            pb = rows[:, blank_id]
            if is_log:
                pb = pb.exp()
            pb = pb.to(torch.float32)
            s = 0
            for e in range(1, L + 1):
                if e == L or ids[e] != ids[s]:
                    if ids[s] == blank_id:
                        for t in range(s, e):
                            segs.append((t, 1))
                            scores.append(pb[t])
                    else:
                        segs.append((s, e - s))
                        scores.append(pb[s:e].mean())
                    s = e
        if scores:
            sc = torch.stack(scores).to(torch.float32)
            keep = (sc < blank_threshold).numpy()
        else:
            sc = torch.zeros(0)
            keep = np.zeros(0, dtype=bool)
        out.append({
            "segs": [sg for sg, k in zip(segs, keep) if k],
            "all_segs": segs,
            "scores": sc.numpy(),
            "keep": keep,
        })
    return ids_all, out, is_log


def psd_loop(encoder_out: torch.Tensor, lens, ctc_posterior: torch.Tensor,
             blank_id: int = 0, blank_threshold: float = BLANK_THRESHOLD):
    """Full PSD (plan + segmented mean-pool + zero pad) — ps-slm.py:237-317.
    Features are pooled from ``encoder_out`` as given (not exp'd in log mode),
    pad rows are zeros (:308-314), ``new_lens`` is int64 (:315).  Edge cases:
    L==0 (:261-264) and all-empty (:304-306)."""
    B, T, D = encoder_out.shape
    _, plan, _ = psd_plan(ctc_posterior, lens, blank_id, blank_threshold)
    new_lens = [len(p["segs"]) for p in plan]
    m = max(new_lens) if new_lens else 0
    if m == 0:
        return encoder_out.new_zeros(B, 0, D), torch.zeros(B, dtype=torch.long)
    out = encoder_out.new_zeros(B, m, D)
    for b, p in enumerate(plan):
        for j, (s, n) in enumerate(p["segs"]):
            out[b, j] = encoder_out[b, s] if n == 1 else encoder_out[b, s:s + n].mean(dim=0)
    return out, torch.tensor(new_lens, dtype=torch.long)


def psd_vec(encoder_out: torch.Tensor, lens: torch.Tensor, ctc_posterior: torch.Tensor,
            blank_id: int = 0, blank_threshold: float = BLANK_THRESHOLD):
    """Vectorised PSD for benchmark-size parity (same semantics as psd_loop;
    proven equal to it and to the reference on random small cases in tests).
    Returns (padded [B,maxM,D], new_lens int64, dict(plan arrays))."""
    B, T, D = encoder_out.shape
    lens = torch.as_tensor(lens, dtype=torch.long)
    is_log = bool(ctc_posterior.max() <= 0) if ctc_posterior.numel() else False
    ids = ctc_posterior.argmax(dim=-1)                                    # [B,T]
    pb = ctc_posterior[..., blank_id].to(torch.float32)
    if is_log:
        pb = pb.exp()
    t_idx = torch.arange(T).unsqueeze(0).expand(B, T)
    valid = t_idx < lens.unsqueeze(1)
    prev = torch.cat([torch.full((B, 1), -1, dtype=ids.dtype), ids[:, :-1]], dim=1)
    start = valid & ((t_idx == 0) | (ids != prev) | (ids == blank_id))
    # global candidate index of every valid frame
    flat_start = start.reshape(-1)
    cand = torch.cumsum(flat_start.to(torch.long), 0) - 1                 # [B*T]
    n_cand = int(flat_start.sum())
    flat_valid = valid.reshape(-1)
    cand_v = cand[flat_valid]
    seg_len = torch.zeros(n_cand, dtype=torch.long).index_add_(0, cand_v, torch.ones_like(cand_v))
    seg_sum = torch.zeros(n_cand, dtype=torch.float32).index_add_(0, cand_v, pb.reshape(-1)[flat_valid])
    score = seg_sum / seg_len.to(torch.float32)
    keep = score < blank_threshold
    cand_b = (torch.nonzero(flat_start).squeeze(-1) // T)
    cand_t = (torch.nonzero(flat_start).squeeze(-1) % T)
    new_lens = torch.zeros(B, dtype=torch.long).index_add_(0, cand_b[keep], torch.ones(int(keep.sum()), dtype=torch.long))
    m = int(new_lens.max()) if B else 0
    plan = {"ids": ids, "cand_b": cand_b, "cand_t": cand_t, "cand_len": seg_len,
            "score": score, "keep": keep, "is_log": is_log}
    if m == 0:
        return encoder_out.new_zeros(B, 0, D), torch.zeros(B, dtype=torch.long), plan
    # kept-candidate rank inside its utterance
    kept_idx = torch.nonzero(keep).squeeze(-1)
    kb = cand_b[kept_idx]
    first_of_b = torch.zeros(B + 1, dtype=torch.long)
    first_of_b[1:] = torch.cumsum(new_lens, 0)
    rank = torch.arange(kept_idx.numel()) - first_of_b[kb]
    # segmented mean over kept candidates only
    kept_of_cand = torch.full((n_cand,), -1, dtype=torch.long)
    kept_of_cand[kept_idx] = torch.arange(kept_idx.numel())
    frame_k = kept_of_cand[cand_v]                                        # per valid frame
    sel = frame_k >= 0
    feats = encoder_out.reshape(B * T, D)[flat_valid][sel]
    pooled = torch.zeros(kept_idx.numel(), D, dtype=encoder_out.dtype).index_add_(0, frame_k[sel], feats)
    klen = seg_len[kept_idx]
    multi = klen > 1
    pooled[multi] = pooled[multi] / klen[multi].to(encoder_out.dtype).unsqueeze(1)
    out = encoder_out.new_zeros(B, m, D)
    out[kb, rank] = pooled
    plan.update({"kept_b": kb, "kept_t": cand_t[kept_idx], "kept_len": klen})
    return out, new_lens, plan


# --------------------------------------------------------------------------
# (a3) clean simulator — Multitask/model/ps-slm.py:337-358
# --------------------------------------------------------------------------
def sim_posterior_clean(ids_list: List[List[int]], vocab_size: int):
    """One-hot [B, L_max, V] fp32 + lens int64 (CPU) — ps-slm.py:346-358."""
    lens = torch.tensor([len(i) for i in ids_list], dtype=torch.long)
    lmax = int(lens.max())
    post = torch.zeros(len(ids_list), lmax, vocab_size, dtype=torch.float32)
    for b, ids in enumerate(ids_list):
        for t, v in enumerate(ids):
            post[b, t, v] = 1.0
    return post, lens


# --------------------------------------------------------------------------
# (a4) noisy simulator — Multitask/model/ps-slm.py:360-409
# --------------------------------------------------------------------------
def sim_noise_decisions(ids_list: List[List[int]], vocab_size: int, blank_id: int = 0,
                        drop_prob: float = 0.05, insert_prob: float = 0.0,
                        smooth_low: float = 0.0, smooth_high: float = 0.1):
    """Draw the simulator's random decisions with torch's CPU global generator
    in the reference's order (per utterance: one uniform_ for alpha :384, then
    rand(L) :387, then per insert randint(0,len+1) + rand(1) :391-392).

    Returns per utterance ``(alpha, rows)`` where rows is a list of
    ``(token_id, is_hard_blank)``: soft rows are (1-alpha)*onehot + alpha/V
    (:385), hard rows are the inserted exact one-hot blank (:397-399); a
    duplicate insert copies its left neighbour (or row 0) whatever it is (:394)."""
    out = []
    for ids in ids_list:
        alpha = torch.empty(()).uniform_(smooth_low, smooth_high).item()
        keep = torch.rand(len(ids)) > drop_prob
        rows = [(int(v), False) for v, k in zip(ids, keep.tolist()) if k]
        n_insert = int(len(rows) * insert_prob)
        for _ in range(n_insert):
            pos = torch.randint(0, len(rows) + 1, (1,)).item()
            if torch.rand(1) < 0.5 and len(rows) > 0:
                dup = rows[pos - 1] if pos > 0 else rows[0]
                rows.insert(pos, dup)
            else:
                rows.insert(pos, (blank_id, True))
        out.append((alpha, rows))
    return out


def sim_row_values(alpha: float, vocab_size: int) -> Tuple[np.float32, np.float32]:
    """fp32 values of a soft row, computed the way torch does at ps-slm.py:385:
    ``(1 - alpha) * onehot + alpha / V`` with python-double scalars applied to
    an fp32 tensor → hot = fl32(fl32(1-alpha)*1 + fl32(alpha/V)), base = fl32(alpha/V)."""
    a = np.float32(1.0 - alpha)
    c = np.float32(alpha / vocab_size)
    return np.float32(a + c), c


def sim_posterior_noise(ids_list, vocab_size, blank_id=0, **kw):
    """Dense [B, L_max, V] fp32 + lens (ps-slm.py:403-409) from fresh decisions."""
    dec = sim_noise_decisions(ids_list, vocab_size, blank_id, **kw)
    return sim_rows_to_dense(dec, vocab_size, blank_id)


def sim_rows_to_dense(dec, vocab_size, blank_id=0):
    lens = torch.tensor([len(r) for _, r in dec], dtype=torch.long)
    lmax = int(lens.max()) if len(dec) else 0
    post = torch.zeros(len(dec), lmax, vocab_size, dtype=torch.float32)
    for b, (alpha, rows) in enumerate(dec):
        hot, base = sim_row_values(alpha, vocab_size)
        for t, (v, hard) in enumerate(rows):
            if hard:
                post[b, t, blank_id] = 1.0
            else:
                post[b, t, :] = float(base)
                post[b, t, v] = float(hot)
    return post, lens


# --------------------------------------------------------------------------
# (a5)/(a6) projectors — Multitask/model/projector.py
# --------------------------------------------------------------------------
def projector_linear_silu(x, norm_w, norm_b, w1, b1, w2, b2, eps=1e-5):
    """LayerNorm(V) → Linear → SiLU → Linear — projector.py:139-151."""
    h = F.layer_norm(x, (x.shape[-1],), norm_w, norm_b, eps)
    return F.linear(F.silu(F.linear(h, w1, b1)), w2, b2)


def _downsample(x, k):
    """drop trailing T % k frames, view [B, T//k, D*k] — projector.py:19-24, :40-46."""
    B, T, D = x.shape
    T2 = (T // k) * k
    return x[:, :T2, :].contiguous().view(B, T2 // k, D * k)


def projector_concat(x, k, w1, b1, w2, b2):
    """k-concat → Linear → ReLU → Linear — projector.py:39-50 ("linear")."""
    return F.linear(F.relu(F.linear(_downsample(x, k), w1, b1)), w2, b2)


def projector_linear(x, k, w, b):
    """k-concat → Linear — projector.py:18-26 ("simple_linear")."""
    return F.linear(_downsample(x, k), w, b)


def projector_ctcca(post, llm_embed, wq, n_heads: int = 8):
    """Cross-attention projector ``EncoderProjectorCTCCA.forward`` — projector.py:104-126: Q = W_q·post, multi-head
    softmax attention of Q over the LLM embedding table (keys = values = table), heads concatenated."""
    B, T, _ = post.shape
    Q = F.linear(post, wq)                                         # :113
    h = n_heads
    d = Q.size(-1) // h
    q = Q.view(B, T, h, d)
    k = llm_embed.view(-1, h, d)
    scores = torch.einsum("bthd,vhd->bthv", q, k) / d ** 0.5        # :121
    attn = scores.softmax(dim=-1)                                  # :122
    z = torch.einsum("bthv,vhd->bthd", attn, k)                    # :123
    return z.contiguous().view(B, T, -1)                           # :124


def voca_trans(encoder_out, encoder_out_lens, w_map, b_map, k, embed_matrix, do_psd, top1_emb, blank_id=151643):
    """Vocabulary-transfer branch — Multitask/model/ps-slm.py:485-516 (forward) / :615-646 (generate).  The shipped
    reference reads ``encoder_outs`` / ``encoder_feature_length`` there before assigning them (UnboundLocalError); this
    restates the branch with ``encoder_out`` / ``encoder_out_lens``, the only candidates in scope, and is pinned against
    the reference source with exactly that one-line fix applied in memory (oracle/make_golden.py, golden case
    ``infer_voca_trans``)."""
    logits = projector_linear(encoder_out, k, w_map, b_map)             # :488  simple_linear as a CTC head
    lens = encoder_out_lens // k                                        # :489
    if do_psd:
        post = torch.softmax(logits, dim=-1)                            # :490
        outs, lens, _ = psd_vec(logits, lens, post, blank_id, 0.9)      # :491  features = the logits themselves
        v_real = outs.size(-1) - 1                                      # :494
        ctc = torch.softmax(outs[..., :v_real], dim=-1)                 # :495-496
        res = torch.einsum("btv,vh->bth", ctc, embed_matrix[:v_real])   # :497
    else:
        ctc = torch.softmax(logits, dim=-1)                             # :510
        res = torch.einsum("btv,vh->bth", ctc, embed_matrix[:logits.size(-1)])   # :511
    if top1_emb:
        res = embed_matrix[ctc.argmax(dim=-1)]                          # :500-502, :514-516
    return res, lens


# --------------------------------------------------------------------------
# (a8) splice — Multitask/model/ps-slm.py:679-873
# --------------------------------------------------------------------------
def merge(audio_features: torch.Tensor, num_audio_tokens, inputs_embeds: torch.Tensor,
          input_ids, attention_mask, labels, speech_id: int, pad_id: int, ignore: int = -100):
    """Restatement of _merge_input_ids_with_audio_features with integer numpy
    planning and explicit per-position placement.  Returns the same 5-tuple
    (emb, mask, labels|None, position_ids, final_input_ids); raises ValueError
    on both-side padding (:783-785) and on audio-slot count mismatch (:861-865)."""
    ids = np.asarray(input_ids, dtype=np.int64)
    att = np.asarray(attention_mask)
    att_dtype = attention_mask.dtype if isinstance(attention_mask, torch.Tensor) else torch.long
    M = np.asarray(num_audio_tokens, dtype=np.int64)
    B, S = ids.shape
    H = inputs_embeds.shape[-1]
    # padding side (:771-785)
    any_left = bool((att[:, 0] == 0).any())
    any_right = bool((att[:, -1] == 0).any())
    left = True
    if B > 1:
        if any_left and any_right:
            raise ValueError("both side of attention_mask has zero, invalid.")
        left = not (not any_left and any_right)
    sp = ids == speech_id
    if int(sp.sum()) != M.shape[0] and M.shape[0] != 1:
        raise IndexError("number of speech tokens does not match number of audios")
    ph = np.ones((B, S), dtype=np.int64)
    ph[sp] = M if M.shape[0] != 1 else M[0]                                 # :805-807 (row-major assignment)
    new_pos = np.cumsum(ph, axis=1) - 1                                      # :808
    tot = ph.sum(axis=1)
    S2 = int(tot.max())                                                      # :809
    shift = (S2 - 1 - new_pos[:, -1]) if left else np.zeros(B, dtype=np.int64)  # :810-812
    n_pad = (att == 0).sum(axis=1)
    span = tot - n_pad

    emb = torch.zeros(B, S2, H, dtype=inputs_embeds.dtype)
    fmask = np.zeros((B, S2), dtype=np.int64)
    fids = np.full((B, S2), pad_id, dtype=np.int64)
    flab = np.full((B, S2), ignore, dtype=np.int64)
    lab = None if labels is None else np.asarray(labels, dtype=np.int64)
    written = np.zeros((B, S2), dtype=bool)
    for b in range(B):
        for j in range(S):
            if (not sp[b, j]) and att[b, j] == 1:                            # :797-799
                p = new_pos[b, j] + shift[b]
                emb[b, p] = inputs_embeds[b, j]                              # :833-835
                fmask[b, p] = att[b, j]
                fids[b, p] = ids[b, j]
                if lab is not None:
                    flab[b, p] = lab[b, j]
                written[b, p] = True
    pidx = np.arange(S2)[None, :]
    if left:
        val = (S2 - pidx) <= span[:, None]                                   # :848-853
    else:
        val = pidx < span[:, None]                                           # :856
    slots = (~written) & val
    # packed audio rows, row-major over (b, t < M_b) — :765-769
    na, maxa = audio_features.shape[0], audio_features.shape[1]
    amask = np.arange(maxa)[None, :] < (M[:, None] if M.shape[0] == na else np.broadcast_to(M, (na,))[:, None])
    packed = audio_features[torch.from_numpy(amask)]
    if int(slots.sum()) != int(M.sum()):                                     # :861-865
        raise ValueError("The input provided to the model are wrong. audio slot count mismatch")
    emb[torch.from_numpy(slots)] = packed                                    # :867-869 (row-major order)
    fmask = fmask | slots                                                    # :870
    pos_ids = np.cumsum(fmask, axis=1) - 1
    pos_ids[fmask == 0] = 1                                                  # :871
    out_labels = None if lab is None else torch.from_numpy(flab)
    return (emb, torch.from_numpy(fmask).to(att_dtype), out_labels,
            torch.from_numpy(pos_ids), torch.from_numpy(fids))


# --------------------------------------------------------------------------
# whole bridge (inference dispatch) — ps-slm.py:581-658 with the default flags
# ctc_posterior=True, voca_trans=False, gt_emb=False, do_psd=True
# --------------------------------------------------------------------------
def bridge_inference(raw_encoder_out, raw_lens, w_ctc, b_ctc, proj_params, embed_table,
                     input_ids, attention_mask, labels, speech_id, pad_id, blank_id=0,
                     vectorised=True):
    post, lens = ctc_head_posterior(raw_encoder_out, raw_lens, w_ctc, b_ctc)
    if vectorised:
        feats, new_lens, _ = psd_vec(post, lens, post, blank_id)
    else:
        feats, new_lens = psd_loop(post, lens, post, blank_id)
    proj = projector_linear_silu(feats, *proj_params)
    text = F.embedding(input_ids, embed_table)
    return merge(proj, new_lens, text, input_ids, attention_mask, labels, speech_id, pad_id), new_lens
