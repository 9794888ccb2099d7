"""Synthetic workloads for tests and bench (no dataset, tokenizer asset or checkpoint is
reachable offline).  Shapes and constants follow BASELINE.md §2 / SURVEY.md §8(d):
V=25055 CTC classes, 512-d encoder output, 60 ms frames, 4 prefix query frames,
Qwen2.5-1.5B width 1536 with 151936 embedding rows.

Planted-label generator: every utterance gets a label track (tokens arrive as a Bernoulli
process, each held 1–3 frames, blank elsewhere); the encoder output of a frame is its label's
unit CTC-weight row scaled so the softmax puts a chosen probability on the label, plus noise.
Token frames and 90 % of the blank frames are confident (p ≥ 0.99), 10 % of the blank frames are
"soft" (p_blank ∈ [0.5, 0.85], kept by PSD) and nothing sits within ±0.01 of the 0.9 threshold,
so integer results are insensitive to bf16 rounding of the GEMM operands.
"""
import math
from typing import List, Optional

import torch

V_CTC = 25055
D_ENC = 512
H_LLM = 1536
V_LLM = 151936
N_PREFIX = 4
SPEECH_ID = 151665      # Qwen2.5 id of the first added special token (parameter, not evidenced in the repo)
PAD_ID = 151643         # eos used as pad (ps-slm.py:27)
IGNORE_ID = -100

# synthetic stand-ins for prompt lengths incl. chat template (ASR, EN2ZH, EN2DE, QA, SLU_scenario)
TASK_PROMPT_LENS = {"ASR": 22, "EN2ZH": 20, "EN2DE": 20, "QA": 17, "SLU_scenario": 85}


def _scale_for_prob(p: torch.Tensor, V: int, d: int) -> torch.Tensor:
    """logit scale s such that softmax puts ≈p on the planted label against V-1 competitors whose
    logits are N(0, s²/d): s = logit(p) + log(V-1) + s²/(2d) (three fixed-point steps)."""
    base = torch.log(p / (1 - p)) + math.log(V - 1)
    s = base.clone()
    for _ in range(4):
        s = base + s * s / (2 * d)
    return s


def make_ctc_head(V: int = V_CTC, d: int = D_ENC, seed: int = 2025):
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(V, d, generator=g) / math.sqrt(d)
    b = torch.zeros(V)
    return w, b


def make_label_tracks(B: int, T: int, V: int, g: torch.Generator, rate: float = 0.21):
    """[B,T] int64 label track: 0 = blank."""
    lab = torch.zeros(B, T, dtype=torch.long)
    arrive = torch.rand(B, T, generator=g) < rate
    dur = torch.multinomial(torch.tensor([0.6, 0.3, 0.1]), B * T, replacement=True, generator=g).view(B, T) + 1
    tok = torch.randint(1, V, (B, T), generator=g)
    for b in range(B):
        t = 0
        a, d_, k = arrive[b].tolist(), dur[b].tolist(), tok[b].tolist()
        row = [0] * T
        while t < T:
            if a[t]:
                n = d_[t]
                for u in range(t, min(T, t + n)):
                    row[u] = k[t]
                t += n
            else:
                t += 1
        lab[b] = torch.tensor(row)
    return lab


def make_encoder_batch(B: int, T: int, w_ctc: torch.Tensor, seed: int = 2025, ragged: bool = False,
                       noise: float = 0.3, return_soft: bool = False):
    """raw_encoder_out [B, T+4, d] fp32, raw_lens [B] int64 (incl. the 4 prefix frames), labels [B,T]
    (+ the soft-blank mask [B,T] with ``return_soft``)."""
    V, d = w_ctc.shape
    g = torch.Generator().manual_seed(seed)
    lab = make_label_tracks(B, T, V, g)
    what = w_ctc / w_ctc.norm(dim=1, keepdim=True)
    wn = w_ctc.norm(dim=1)
    is_blank = lab == 0
    soft = is_blank & (torch.rand(B, T, generator=g) < 0.10)
    p = torch.full((B, T), 0.995)
    p[soft] = 0.5 + 0.35 * torch.rand(int(soft.sum()), generator=g)
    s = _scale_for_prob(p, V, d) / wn[lab]
    x = s.unsqueeze(-1) * what[lab] + noise * torch.randn(B, T, d, generator=g) / math.sqrt(d)
    prefix = torch.randn(B, N_PREFIX, d, generator=g) / math.sqrt(d)
    raw = torch.cat([prefix, x], dim=1).contiguous()
    if ragged:
        lens = torch.randint(T // 4, T + 1, (B,), generator=g)
        lens[0] = T
    else:
        lens = torch.full((B,), T, dtype=torch.long)
    if return_soft:
        return raw, (lens + N_PREFIX).to(torch.long), lab, soft
    return raw, (lens + N_PREFIX).to(torch.long), lab


def expected_plan(lab: torch.Tensor, soft: torch.Tensor, lens: torch.Tensor):
    """What PSD must produce for a planted batch, from first principles (ps-slm.py:268-297): token runs
    merge and are always kept (their blank probability is ~0), blank frames stand alone and are kept
    iff they are soft (p_blank < 0.9).  Returns per utterance the list of (start, length)."""
    out = []
    for b in range(lab.shape[0]):
        L = int(lens[b])
        row, sf = lab[b, :L].tolist(), soft[b, :L].tolist()
        segs, s = [], 0
        for e in range(1, L + 1):
            if e == L or row[e] != row[s]:
                if row[s] == 0:
                    segs.extend((t, 1) for t in range(s, e) if sf[t])
                else:
                    segs.append((s, e - s))
                s = e
        out.append(segs)
    return out


def make_prompts(B: int, seed: int = 2025, tasks: Optional[List[str]] = None, left_pad: bool = True,
                 target_lens: Optional[List[int]] = None):
    """input_ids [B,S] int64, attention_mask [B,S] bool, labels [B,S] int64|None.
    One <speech> token per row placed where the chat template puts it (3 tokens before the end of
    the prompt); inference rows are left-padded prompts, training rows are right-padded
    prompt+target+eos with -100 on the prompt (speech_dataset_large.py:162-186, :243-264)."""
    g = torch.Generator().manual_seed(seed + 17)
    names = list(TASK_PROMPT_LENS)
    rows, labs = [], []
    for b in range(B):
        task = tasks[b % len(tasks)] if tasks else names[int(torch.randint(0, len(names), (1,), generator=g))]
        n = TASK_PROMPT_LENS[task]
        ids = torch.randint(0, PAD_ID, (n,), generator=g)
        ids[n - 6] = SPEECH_ID
        if target_lens is not None:
            tgt = torch.randint(0, PAD_ID, (target_lens[b],), generator=g)
            row = torch.cat([ids, tgt, torch.tensor([PAD_ID])])
            lab = torch.cat([torch.full((n,), IGNORE_ID), tgt, torch.tensor([PAD_ID])])
        else:
            row, lab = ids, None
        rows.append(row)
        labs.append(lab)
    S = max(len(r) for r in rows)
    input_ids = torch.full((B, S), PAD_ID, dtype=torch.long)
    mask = torch.zeros(B, S, dtype=torch.bool)
    labels = torch.full((B, S), IGNORE_ID, dtype=torch.long) if target_lens is not None else None
    for b, r in enumerate(rows):
        n = len(r)
        sl = slice(S - n, S) if left_pad else slice(0, n)
        input_ids[b, sl] = r
        mask[b, sl] = True
        if labels is not None:
            labels[b, sl] = labs[b]
    return input_ids, mask, labels


def make_transcripts(B: int, V: int = V_CTC, seed: int = 1234, lo: int = 35, hi: int = 105, repeat: float = 0.05):
    """Synthetic GT token-id lists (10–30 s at ≈3.5 tokens/s) with a few adjacent repeats."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(B):
        n = int(torch.randint(lo, hi + 1, (1,), generator=g))
        ids = torch.randint(1, V, (n,), generator=g).tolist()
        r = torch.rand(n, generator=g).tolist()
        for t in range(1, n):
            if r[t] < repeat:
                ids[t] = ids[t - 1]
        out.append(ids)
    return out


def make_embed_table(rows: int = V_LLM, H: int = H_LLM, dtype=torch.bfloat16, seed: int = 0, device="cpu"):
    g = torch.Generator(device=device).manual_seed(seed)
    return (torch.randn(rows, H, generator=g, device=device) * 0.02).to(dtype)
