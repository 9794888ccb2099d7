"""Text-simulated CTC posteriors (step 1a of the bridge), device-side construction.

* ``ctc_pseudo_posterior``        ← slam_model_asr.ctc_pseudo_posterior        (ps-slm.py:337-358)
* ``ctc_pseudo_posterior_noise``  ← slam_model_asr.ctc_pseudo_posterior_noise  (ps-slm.py:360-409)

The reference builds ``[B, L_max, 25055]`` fp32 on the CPU (one_hot, boolean gathers, O(n_insert)
``torch.cat`` reallocations) and then copies ~100 KB per token over PCIe.  Here only the random
DECISIONS are made on the host — drawn from torch's CPU global generator in exactly the
reference's order so results are reproducible bit for bit — and a few bytes per row go to the
device, where one kernel writes the rows at HBM speed (or, for the fused training path, writes
bf16 rows plus closed-form LayerNorm statistics).
"""
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops


def draw_noise_decisions(ids_list: Sequence[Sequence[int]], blank_id: int = 0, drop_prob: float = 0.05,
                         insert_prob: float = 0.0, smooth_low: float = 0.0, smooth_high: float = 0.1):
    """Per utterance ``(alpha, rows)``, rows = [(token_id, is_hard_blank)].  RNG order of the
    reference: one ``uniform_`` (:384), ``rand(L)`` (:387), then per insert ``randint(0, len+1)`` and
    ``rand(1)`` (:391-392)."""
    out = []
    for ids in ids_list:
        alpha = torch.empty(()).uniform_(smooth_low, smooth_high).item()
        keep = (torch.rand(len(ids)) > drop_prob).tolist()
        rows = [(int(v), False) for v, k in zip(ids, keep) if k]
        for _ in range(int(len(rows) * insert_prob)):
            pos = torch.randint(0, len(rows) + 1, (1,)).item()
            if torch.rand(1) < 0.5 and len(rows) > 0:
                rows.insert(pos, rows[pos - 1] if pos > 0 else rows[0])     # duplicate the left neighbour (:394)
            else:
                rows.insert(pos, (blank_id, True))                          # exact one-hot blank (:397-399)
        out.append((alpha, rows))
    return out


def draw_noise_descriptors(ids_list: Sequence[Sequence[int]], vocab_size: int, blank_id: int = 0,
                           drop_prob: float = 0.05, insert_prob: float = 0.0, smooth_low: float = 0.0,
                           smooth_high: float = 0.1):
    """Same random stream as ``draw_noise_decisions`` but straight to flat numpy row descriptors
    ``(tok int32, hot f32, base f32, lens)`` without per-token Python objects (the host side of a training step
    must not dominate it).  Falls back to the list algorithm only for utterances that get inserts."""
    if int(max((len(i) for i in ids_list), default=0) * insert_prob) == 0:
        return _draw_noise_descriptors_batched(ids_list, vocab_size, drop_prob, smooth_low, smooth_high)
    toks, hots, bases, lens = [], [], [], []
    for ids in ids_list:
        alpha = torch.empty(()).uniform_(smooth_low, smooth_high).item()
        keep = (torch.rand(len(ids)) > drop_prob).numpy()
        tok = np.asarray(ids, dtype=np.int32)[keep]
        h, c = soft_row_values(alpha, vocab_size)
        n_insert = int(tok.shape[0] * insert_prob)
        if n_insert == 0:
            hot = np.full(tok.shape[0], h, dtype=np.float32)
            base = np.full(tok.shape[0], c, dtype=np.float32)
        else:
            rows = [(int(v), False) for v in tok]
            for _ in range(n_insert):
                pos = torch.randint(0, len(rows) + 1, (1,)).item()
                if torch.rand(1) < 0.5 and len(rows) > 0:
                    rows.insert(pos, rows[pos - 1] if pos > 0 else rows[0])
                else:
                    rows.insert(pos, (blank_id, True))
            tok = np.asarray([r[0] for r in rows], dtype=np.int32)
            hard = np.asarray([r[1] for r in rows], dtype=bool)
            hot = np.where(hard, np.float32(1.0), h).astype(np.float32)
            base = np.where(hard, np.float32(0.0), c).astype(np.float32)
        toks.append(tok); hots.append(hot); bases.append(base); lens.append(int(tok.shape[0]))
    cat = (lambda xs, dt: np.concatenate(xs) if xs else np.zeros(0, dtype=dt))
    return cat(toks, np.int32), cat(hots, np.float32), cat(bases, np.float32), lens


def _draw_noise_descriptors_batched(ids_list, vocab_size: int, drop_prob: float, smooth_low: float, smooth_high: float):
    """No-insert case (the shipped default, insert_prob=0): ONE ``torch.rand`` call for the whole batch.
    torch's CPU generator hands out one 24-bit draw per fp32 uniform in call order, so ``rand(sum(L_b + 1))`` is
    the same stream as the reference's per-utterance ``uniform_()`` (:384) + ``rand(L_b)`` (:387) sequence
    (checked bit for bit in tests/test_oracle_vs_reference.py); ``uniform_(lo, hi)`` is ``u·(hi − lo) + lo`` in fp32."""
    lens_in = np.fromiter((len(i) for i in ids_list), dtype=np.int64, count=len(ids_list))
    B = lens_in.shape[0]
    total = int(lens_in.sum()) + B
    u = torch.rand(total).numpy()
    starts = np.zeros(B + 1, dtype=np.int64)
    np.cumsum(lens_in + 1, out=starts[1:])
    lo, hi = np.float32(smooth_low), np.float32(smooth_high)
    alpha = (u[starts[:-1]] * (hi - lo) + lo).astype(np.float32).astype(np.float64)      # .item() → python float
    is_alpha = np.zeros(total, dtype=bool)
    is_alpha[starts[:-1]] = True
    keep = u[~is_alpha] > np.float32(drop_prob)
    tok_all = (np.concatenate([np.asarray(i, dtype=np.int32) for i in ids_list]) if total > B
               else np.zeros(0, dtype=np.int32))
    utt = np.repeat(np.arange(B), lens_in)
    tok = tok_all[keep]
    utt_k = utt[keep]
    lens = np.bincount(utt_k, minlength=B).astype(np.int64) if B else np.zeros(0, dtype=np.int64)
    a32 = (1.0 - alpha).astype(np.float32)                      # python double → fp32 scalar, as torch does (:385)
    c32 = (alpha / vocab_size).astype(np.float32)
    h32 = (a32 + c32).astype(np.float32)
    return tok, h32[utt_k], c32[utt_k], [int(x) for x in lens]


def clean_descriptors(ids_list: Sequence[Sequence[int]]):
    """Descriptors of the clean one-hot simulator (ps-slm.py:337-358): hot = 1, base = 0."""
    lens = [len(i) for i in ids_list]
    n = sum(lens)
    tok = (np.concatenate([np.asarray(i, dtype=np.int32) for i in ids_list]) if n else np.zeros(0, dtype=np.int32))
    return tok, np.ones(n, dtype=np.float32), np.zeros(n, dtype=np.float32), lens


def soft_row_values(alpha: float, vocab_size: int) -> Tuple[np.float32, np.float32]:
    """fp32 (hot, base) of ``(1 - alpha) * onehot + alpha / V`` as torch evaluates it (:385)."""
    a = np.float32(1.0 - alpha)
    c = np.float32(alpha / vocab_size)
    return np.float32(a + c), c


def _descriptors(decisions, vocab_size: int, pad_to_max: bool):
    lens = [len(r) for _, r in decisions]
    lmax = max(lens) if lens else 0
    tok, hot, base = [], [], []
    for (alpha, rows), n in zip(decisions, lens):
        h, c = soft_row_values(alpha, vocab_size)
        for v, hard in rows:
            tok.append(v)
            hot.append(1.0 if hard else float(h))
            base.append(0.0 if hard else float(c))
        if pad_to_max:
            tok.extend([-1] * (lmax - n)); hot.extend([0.0] * (lmax - n)); base.extend([0.0] * (lmax - n))
    return (np.asarray(tok, dtype=np.int32), np.asarray(hot, dtype=np.float32), np.asarray(base, dtype=np.float32),
            lens, lmax)


def _to_dev(a: np.ndarray, device):
    t = torch.from_numpy(a)
    if t.numel():
        t = t.pin_memory()
    return t.to(device, non_blocking=True)


def build_dense(decisions, vocab_size: int, device, dtype=torch.float32):
    """[B, L_max, V] posterior + lens (int64, on device) from decisions — the reference's return contract."""
    tok, hot, base, lens, lmax = _descriptors(decisions, vocab_size, True)
    B = len(decisions)
    out = torch.empty(B, lmax, vocab_size, dtype=dtype, device=device)
    if B * lmax:
        ops.sim_posterior_rows(_to_dev(tok, device), _to_dev(hot, device), _to_dev(base, device), vocab_size,
                               out, vocab_size)
    return out, torch.tensor(lens, dtype=torch.long, device=device)


def build_packed_bf16(decisions, vocab_size: int, device, ln_eps: float = 1e-5):
    """Packed bf16 rows [sum L_b, pad64(V)] + LayerNorm stats + lens: the A operand of the folded
    projector GEMM for the text-only training step (never materialises the fp32 posterior).
    ``decisions`` is either the list of ``draw_noise_decisions`` or the tuple of ``draw_noise_descriptors``."""
    if isinstance(decisions, tuple):
        tok, hot, base, lens = decisions
    else:
        tok, hot, base, lens, _ = _descriptors(decisions, vocab_size, False)
    n = int(tok.shape[0])
    ld = ops.pad_to(vocab_size)
    rows = torch.empty(n, ld, dtype=torch.bfloat16, device=device)
    mean = torch.empty(n, dtype=torch.float32, device=device)
    rstd = torch.empty(n, dtype=torch.float32, device=device)
    if n:
        ops.sim_posterior_rows(_to_dev(tok, device), _to_dev(hot, device), _to_dev(base, device), vocab_size,
                               rows, ld, ln_mean=mean, ln_rstd=rstd, ln_eps=ln_eps)
    return rows, mean, rstd, torch.tensor(lens, dtype=torch.long, device=device)


def ctc_pseudo_posterior(ids_list: List[List[int]], vocab_size: int, device):
    """One-hot posterior [B, L_max, V] fp32 and lens (ps-slm.py:337-358).  The reference returns CPU
    tensors which its caller moves to the device (:467-468, :597-598); this returns them there."""
    dec = [(0.0, [(int(v), True) for v in ids]) for ids in ids_list]
    # hard rows: exactly 1.0 at the token, 0 elsewhere (here "hard" just means no smoothing)
    return build_dense(dec, vocab_size, device)


def ctc_pseudo_posterior_noise(ids_list: List[List[int]], vocab_size: int, device, blank_id: int = 0,
                               drop_prob: float = 0.05, insert_prob: float = 0.0, smooth_low: float = 0.0,
                               smooth_high: float = 0.1):
    """Smoothed / dropped / inserted pseudo-posterior and lens on ``device`` (ps-slm.py:360-409)."""
    dec = draw_noise_decisions(ids_list, blank_id, drop_prob, insert_prob, smooth_low, smooth_high)
    return build_dense(dec, vocab_size, device)


class TokenRowPrefetcher:
    """Runs the noisy simulator (ps-slm.py:360-409) one batch AHEAD of the training step on a host thread:
    ``torch.rand`` + the native descriptor call + the pinned host→device copy (on a side stream) of batch i+1
    overlap the kernels and launches of batch i.  Draw order — hence the RNG stream — is the iteration order.

        for rows in TokenRowPrefetcher(batches, V, device):      # batches: iterable of ops.TokenBatch
            y = projector.forward_token_rows(rows)
    ``seeds`` (optional iterable) re-seeds torch's CPU generator before each draw (reproducible benchmarks)."""

    def __init__(self, batches, vocab_size: int, device, depth: int = 2, seeds=None, drop_prob: float = 0.05,
                 smooth_low: float = 0.0, smooth_high: float = 0.1):
        import queue
        import threading
        self.V, self.device = vocab_size, torch.device(device)
        self.kw = dict(drop_prob=drop_prob, smooth_low=smooth_low, smooth_high=smooth_high)
        self.q = queue.Queue(maxsize=depth)
        self.stream = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None
        self.err = None
        self.t = threading.Thread(target=self._run, args=(iter(batches), iter(seeds) if seeds is not None else None),
                                  daemon=True)
        self.t.start()

    def _run(self, it, seeds):
        try:
            if self.stream is not None:
                torch.cuda.set_device(self.device)
            for batch in it:
                if seeds is not None:
                    torch.manual_seed(next(seeds))
                if self.stream is None:
                    self.q.put((ops.sim_token_rows(batch, self.V, self.device, **self.kw), None))
                    continue
                with torch.cuda.stream(self.stream):
                    rows = ops.sim_token_rows(batch, self.V, self.device, **self.kw)
                    ev = torch.cuda.Event()
                    ev.record(self.stream)
                self.q.put((rows, ev))
        except BaseException as e:  # noqa: BLE001 — re-raised in the consumer
            self.err = e
        self.q.put(None)

    def __iter__(self):
        while True:
            item = self.q.get()
            if item is None:
                if self.err is not None:
                    raise self.err
                return
            rows, ev = item
            if ev is not None:
                cur = torch.cuda.current_stream(self.device)
                cur.wait_event(ev)
                rows.hot.record_stream(cur)              # all descriptor tensors are views of one staging copy
            yield rows
