"""Tiny stand-ins for the parts of TASU that are OUT of the bridge's scope (SenseVoice encoder,
Qwen LLM, tokenizers), shared by oracle/make_golden.py (which drives the UNMODIFIED reference
``slam_model_asr.forward/generate`` with them) and by the GPU drop-in test (which drives
``ps_slm_b200.model.slam_model_asr`` with the same objects)."""
import types

import torch
import torch.nn as nn

V, D, H, LLM_VOCAB = 61, 32, 48, 200
SPEECH_ID, PAD_ID, BOS_ID, EOS_ID = 199, 0, 1, 2
VOCA_HEAD, VOCA_LLM_ROWS = 151644, 151650          # CTC head over the LLM vocabulary: 151643 labels + blank (ps-slm.py:491)


class FakeEncoderCore(nn.Module):
    """``self.encoder.encoder(speech, lens)`` → canned planted-label encoder output (+4 prefix frames)."""

    def __init__(self, canned, canned_lens):
        super().__init__()
        self.register_buffer("canned", canned)
        self.register_buffer("canned_lens", canned_lens)

    def forward(self, speech, speech_lengths):
        assert speech.shape[1] == self.canned.shape[1], "fbank frames + 4 query frames"
        return self.canned, self.canned_lens.to(torch.int32)


class FakeSenseVoice(nn.Module):
    blank_id = 0

    def __init__(self, canned, canned_lens, w_ctc, b_ctc):
        super().__init__()
        self.embed = nn.Embedding(16, 560)
        self.encoder = FakeEncoderCore(canned, canned_lens)
        lo = nn.Linear(D, V)
        with torch.no_grad():
            lo.weight.copy_(w_ctc)
            lo.bias.copy_(b_ctc)
        self.ctc = nn.Module()
        self.ctc.ctc_lo = lo


class FakeLLM(nn.Module):
    """Records what the bridge hands to the LLM; returns logits from a linear head."""

    def __init__(self, vocab=LLM_VOCAB):
        super().__init__()
        self.emb = nn.Embedding(vocab, H)
        self.head = nn.Linear(H, LLM_VOCAB)
        self.seen = None

    def get_input_embeddings(self):
        return self.emb

    def forward(self, inputs_embeds=None, attention_mask=None, labels=None, position_ids=None):
        self.seen = dict(inputs_embeds=inputs_embeds, attention_mask=attention_mask, labels=labels,
                         position_ids=position_ids)
        logits = self.head(inputs_embeds.float())
        loss = torch.nn.functional.cross_entropy(logits[:, :-1].reshape(-1, LLM_VOCAB), labels[:, 1:].reshape(-1),
                                                 ignore_index=-100)
        return types.SimpleNamespace(logits=logits, loss=loss)

    def generate(self, inputs_embeds=None, attention_mask=None, **kwargs):
        self.seen = dict(inputs_embeds=inputs_embeds, attention_mask=attention_mask, kwargs=kwargs)
        return self.head(inputs_embeds.float()).argmax(-1)


class FakeLLMTokenizer:
    default_speech_token = SPEECH_ID
    default_ignore_token = -100
    pad_token_id = PAD_ID
    bos_token_id = BOS_ID
    eos_token_id = EOS_ID


class FakeCTCTokenizer:
    vocab_size = V

    def encode(self, text):
        return [int(t) for t in text.split()] if text.strip() else []


class Cfg(dict):
    """attribute + .get access like an OmegaConf DictConfig"""
    __getattr__ = dict.__getitem__


def planted_batch(B, T, seed):
    """Small-vocab version of ps_slm_b200.synth.make_encoder_batch: (canned [B,T+4,D], lens, w, b)."""
    import math
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(V, D, generator=g) / math.sqrt(D)
    b = torch.zeros(V)
    lab = torch.randint(1, V, (B, T), generator=g)
    lab[torch.rand(B, T, generator=g) < 0.6] = 0
    for t in range(1, T):
        rep = torch.rand(B, generator=g) < 0.3
        lab[rep, t] = lab[rep, t - 1]
    what = w / w.norm(dim=1, keepdim=True)
    soft = (lab == 0) & (torch.rand(B, T, generator=g) < 0.25)
    scale = torch.where(soft, torch.tensor(5.5), torch.tensor(14.0)) / w.norm(dim=1)[lab]
    x = scale.unsqueeze(-1) * what[lab] + 0.05 * torch.randn(B, T, D, generator=g) / math.sqrt(D)
    canned = torch.cat([torch.randn(B, 4, D, generator=g) * 0.1, x], 1)
    lens = torch.randint(T // 2, T + 1, (B,), generator=g)
    lens[0] = T
    return canned.contiguous(), lens + 4, w, b


def prompts(B, left_pad, with_labels, seed):
    g = torch.Generator().manual_seed(seed)
    rows, labs = [], []
    for b in range(B):
        n = int(torch.randint(4, 9, (1,), generator=g))
        ids = torch.randint(3, SPEECH_ID, (n,), generator=g)
        ids[int(torch.randint(0, n, (1,), generator=g))] = SPEECH_ID
        if with_labels:
            tl = int(torch.randint(1, 6, (1,), generator=g))
            tgt = torch.randint(3, SPEECH_ID, (tl,), generator=g)
            rows.append(torch.cat([ids, tgt, torch.tensor([EOS_ID])]))
            labs.append(torch.cat([torch.full((n,), -100), tgt, torch.tensor([EOS_ID])]))
        else:
            rows.append(ids)
            labs.append(None)
    S = max(len(r) for r in rows)
    input_ids = torch.full((B, S), PAD_ID, dtype=torch.long)
    mask = torch.zeros(B, S, dtype=torch.bool)
    labels = torch.full((B, S), -100, dtype=torch.long) if with_labels else None
    for b, r in enumerate(rows):
        sl = slice(S - len(r), S) if left_pad else slice(0, len(r))
        input_ids[b, sl] = r
        mask[b, sl] = True
        if with_labels:
            labels[b, sl] = labs[b]
    return input_ids, mask, labels


CASES = {
    # name: (train flags, projector, encoder_dim, ds_rate, left_pad, with_labels, entry)
    "infer_psd": (dict(ctc_posterior=True, do_psd=True, voca_trans=False, gt_emb=False, gt_emb_noise=False, top1_emb=False),
                  "linear-silu", V, 1, True, False, "generate"),
    "train_audio": (dict(ctc_posterior=True, do_psd=True, voca_trans=False, gt_emb=False, gt_emb_noise=False, top1_emb=False),
                    "linear-silu", V, 1, False, True, "forward"),
    "train_text_noise": (dict(ctc_posterior=True, do_psd=True, voca_trans=False, gt_emb=True, gt_emb_noise=True, top1_emb=False),
                         "linear-silu", V, 1, False, True, "forward"),
    "train_text_clean": (dict(ctc_posterior=True, do_psd=True, voca_trans=False, gt_emb=True, gt_emb_noise=False, top1_emb=False),
                         "linear-silu", V, 1, False, True, "forward"),
    "raw_feature_psd": (dict(ctc_posterior=False, do_psd=True, voca_trans=False, gt_emb=False, gt_emb_noise=False, top1_emb=False),
                        "linear", D, 2, True, False, "generate"),
    "posterior_nopsd": (dict(ctc_posterior=True, do_psd=False, voca_trans=False, gt_emb=False, gt_emb_noise=False, top1_emb=False),
                        "linear-silu", V, 1, True, False, "generate"),
    # vocabulary transfer: simple_linear projector = CTC head over the LLM vocabulary, blank id 151643 (ps-slm.py:491)
    "infer_voca_trans": (dict(ctc_posterior=True, do_psd=True, voca_trans=True, gt_emb=False, gt_emb_noise=False, top1_emb=False),
                         "simple_linear", D, 2, True, False, "generate"),
    "infer_cross_attn": (dict(ctc_posterior=True, do_psd=True, voca_trans=False, gt_emb=False, gt_emb_noise=False, top1_emb=False),
                         "cross-attention", V, 1, True, False, "generate"),
}


def build_inputs(name, seed=0):
    flags, proj, enc_dim, k, left, with_labels, entry = CASES[name]
    B, T = 3, 40
    canned, lens, w, b = planted_batch(B, T, 100 + seed)
    input_ids, mask, labels = prompts(B, left, with_labels, 200 + seed)
    g = torch.Generator().manual_seed(300 + seed)
    feats = torch.randn(B, T, 560, generator=g)
    texts = [" ".join(str(int(v)) for v in torch.randint(1, V, (int(torch.randint(3, 9, (1,), generator=g)),), generator=g))
             for _ in range(B)]
    batch = dict(input_ids=input_ids, attention_mask=mask, input_features=feats,
                 input_feature_length=torch.full((B,), T, dtype=torch.long))
    if entry == "forward":
        batch.update(labels=labels, GT=texts)
    else:
        batch.update(targets=texts)
    return dict(flags=flags, proj=proj, enc_dim=enc_dim, k=k, entry=entry, canned=canned, lens=lens, w=w, b=b,
                batch=batch)


def build_parts(inp, projector_cls, seed=0):
    """(encoder, llm, projector, tokenizer, train_config, model_config) with seeded weights."""
    torch.manual_seed(400 + seed)
    encoder = FakeSenseVoice(inp["canned"], inp["lens"], inp["w"], inp["b"])
    voca = bool(inp["flags"].get("voca_trans"))
    llm = FakeLLM(VOCA_LLM_ROWS) if voca else FakeLLM()
    model_config = Cfg(encoder_projector=inp["proj"], encoder_dim=inp["enc_dim"], llm_dim=VOCA_HEAD if voca else H,
                       encoder_projector_ds_rate=inp["k"], encoder_path="unused")
    projector = projector_cls(model_config)
    if voca:
        with torch.no_grad():                          # a peaky head: frames decide for a label or for the blank (151643)
            projector.map.weight.mul_(12.0)
            projector.map.bias.zero_()
            projector.map.bias[VOCA_HEAD - 1] = 3.0
    if inp["proj"] == "linear-silu":
        with torch.no_grad():
            projector.norm.weight.uniform_(0.8, 1.2)
            projector.norm.bias.uniform_(-0.1, 0.1)
    train_config = Cfg(freeze_projector=False, **inp["flags"])
    return encoder, llm, projector, FakeLLMTokenizer(), train_config, model_config
