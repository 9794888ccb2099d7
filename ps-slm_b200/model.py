"""Drop-in plugin surface: ``model_factory`` / ``slam_model_asr`` with the bridge on B200 kernels.

Select it from the reference's launch scripts with
``++model_config.file=<repo>/ps-slm_b200/model.py:model_factory`` — the loader
(Multitask/utils/model_utils.py:9-33) only needs a ``.py`` path and a function name, and calls
``model_factory(train_config, model_config, **kwargs) -> (model, tokenizer)``
(Multitask/finetune_deepspeed.py:127-128, Multitask/inference_batch.py:113-114).

Only the bridge is replaced.  Tokenizer, LLM and SenseVoice encoder are still built by the
reference's own ``setup_tokenizer / setup_llm / setup_encoder`` (Multitask/model/ps-slm.py:25-40,
:89-127), found through ``TASU_REFERENCE_ROOT`` or the current working directory (the reference's
entry points run from ``Multitask/``).  ``slam_model_asr`` below keeps the reference's method
names, signatures, return tuples, flags and printed-nothing behaviour for:

* ``psd``                                   (ps-slm.py:237-317)
* ``ctc_pseudo_posterior`` / ``_noise``     (ps-slm.py:337-409)
* ``_merge_input_ids_with_audio_features``  (ps-slm.py:679-873)
* ``forward`` / ``generate`` dispatch       (ps-slm.py:411-537, :539-677)
"""
import importlib.util
import os
import re
import sys
import types
from typing import List, Optional

import torch
import torch.nn as nn

try:  # imported as part of the package …
    from . import bridge as _bridge, ops as _ops, sim as _sim
    from .projector import PROJECTORS
except ImportError:  # … or loaded by file path through the reference's plugin loader
    _root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if _root not in sys.path:
        sys.path.insert(0, _root)
    import ps_slm_b200.bridge as _bridge
    import ps_slm_b200.sim as _sim
    import ps_slm_b200.ops as _ops
    from ps_slm_b200.projector import PROJECTORS


def _load_reference_module():
    """The reference's model/ps-slm.py (for setup_tokenizer / setup_llm / setup_encoder only)."""
    root = os.environ.get("TASU_REFERENCE_ROOT", os.getcwd())
    path = os.path.join(root, "model", "ps-slm.py")
    if not os.path.isfile(path):
        raise FileNotFoundError(
            "reference tree not found (looked for %s); run from the reference's Multitask/ directory or set "
            "TASU_REFERENCE_ROOT" % path)
    if root not in sys.path:
        sys.path.insert(0, root)
    spec = importlib.util.spec_from_file_location("tasu_reference_ps_slm", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def setup_encoder_projector(train_config, model_config, **kwargs):
    """String → class dispatch of ps-slm.py:43-86 for the in-scope plugins."""
    name = model_config.encoder_projector
    if name not in PROJECTORS:
        raise NotImplementedError(
            "encoder_projector=%r is outside the B200 bridge scope (have: %s)" % (name, sorted(PROJECTORS)))
    encoder_projector = PROJECTORS[name](model_config)
    if name == "linear-silu" and getattr(train_config, "freeze_projector", False):
        for _, param in encoder_projector.named_parameters():
            param.requires_grad = False
        encoder_projector.eval()
    if name == "simple_linear" and getattr(model_config, "ctc_linear", None):
        # pretrained CTC head over the LLM vocabulary (ps-slm.py:67-85): the checkpoint may be wrapped in {"model": ...},
        # the head lives under ctc_head.weight / ctc_head.bias, it is loaded strictly and frozen with the encoder
        ckpt = torch.load(model_config.ctc_linear, map_location="cpu")
        state = ckpt.get("model", ckpt)
        proj_state = {"weight": state["ctc_head.weight"], "bias": state["ctc_head.bias"]}
        encoder_projector.map.load_state_dict(proj_state, strict=True)
        if getattr(train_config, "freeze_encoder", False):
            for _, param in encoder_projector.named_parameters():
                param.requires_grad = False
            encoder_projector.eval()
    return encoder_projector


def model_factory(train_config, model_config, **kwargs):
    ref = _load_reference_module()
    tokenizer = ref.setup_tokenizer(train_config, model_config, **kwargs)
    tokenizer.add_special_tokens({"additional_special_tokens": ["<speech>"]})
    tokenizer.default_ignore_token = -100
    tokenizer.default_speech_token = tokenizer.convert_tokens_to_ids("<speech>")
    llm = ref.setup_llm(train_config, model_config, **kwargs)
    encoder = ref.setup_encoder(train_config, model_config, **kwargs)
    encoder_projector = setup_encoder_projector(train_config, model_config, **kwargs)
    model = slam_model_asr(encoder, llm, encoder_projector, tokenizer, train_config, model_config, **kwargs)
    ckpt_path = kwargs.get("ckpt_path", None)
    if ckpt_path is not None:                       # ps-slm.py:163-170: overlay, strict=False
        model.load_state_dict(torch.load(ckpt_path, map_location="cpu"), strict=False)
    return model, tokenizer


def _cfg_get(cfg, name, default=None):
    if hasattr(cfg, "get"):
        try:
            return cfg.get(name, default)
        except Exception:
            pass
    return getattr(cfg, name, default)


class slam_model_asr(nn.Module):
    def __init__(self, encoder, llm, encoder_projector, tokenizer, train_config, model_config, **kwargs):
        super().__init__()
        self.encoder = encoder
        self.llm = llm
        self.encoder_projector = encoder_projector
        self.tokenizer = tokenizer
        self.metric = kwargs.get("metric", "acc")
        self.train_config = train_config
        self.model_config = model_config
        self.ctc_posterior = _cfg_get(train_config, "ctc_posterior", False)
        self.do_psd = _cfg_get(train_config, "do_psd", False)
        self.voca_trans = _cfg_get(train_config, "voca_trans", False)
        self.gt_emb = _cfg_get(train_config, "gt_emb", False)
        self.gt_emb_noise = _cfg_get(train_config, "gt_emb_noise", False)
        self.top1_emb = _cfg_get(train_config, "top1_emb", False)
        self.cross_attn = model_config.encoder_projector == "cross-attention"
        # voca_trans: the reference's own branch reads `encoder_outs` before assigning it (ps-slm.py:488, :618 —
        # UnboundLocalError as shipped); bridge.voca_trans_project implements it on `encoder_out`, the evident intent
        self.encoder_tokenizer = kwargs.get("encoder_tokenizer", None)
        if self.encoder_tokenizer is None and (self.gt_emb or kwargs.get("need_encoder_tokenizer", False)):
            ref = _load_reference_module()           # SentencePiece wrapper of the reference (host-side, out of scope)
            from model.tokenizer import SenseVoiceTokenizer  # noqa: F401  (reference module on sys.path)
            self.encoder_tokenizer = SenseVoiceTokenizer(model_config.encoder_path)
            del ref
        self._bridge = None
        self._ctc_cache = _bridge.ProjectorCache()
        self._table_cache = _bridge.ProjectorCache()
        # text-only batches go through the token-row projector (no [B, L, 25055] tensor); False = dense simulator path
        self.token_row_path = True

    def train(self, mode: bool = True):
        """Every switch between training and evaluation drops the cached bf16 / folded weight copies
        (deepspeed_utils.py:249-256 evaluates inside the training loop; ZeRO updates parameters through flat buffers
        that move neither a tensor's address nor its version counter)."""
        _bridge.invalidate_caches()
        return super().train(mode)

    # ------------------------------------------------------------------ bridge methods
    def psd(self, encoder_out, encoder_out_lens, ctc_posterior, blank_id: int = 0, blank_threshold: float = 0.90):
        return _bridge.psd(encoder_out, encoder_out_lens, ctc_posterior, blank_id, blank_threshold)

    def _device(self):
        return next(self.parameters()).device

    def ctc_pseudo_posterior(self, texts: List[str]):
        ids_list = [self.encoder_tokenizer.encode(t) for t in texts]
        return _sim.ctc_pseudo_posterior(ids_list, self.encoder_tokenizer.vocab_size, self._device())

    def ctc_pseudo_posterior_noise(self, texts: List[str]):
        ids_list = [self.encoder_tokenizer.encode(t) for t in texts]
        return _sim.ctc_pseudo_posterior_noise(
            ids_list, self.encoder_tokenizer.vocab_size, self._device(), blank_id=self.encoder.blank_id,
            drop_prob=getattr(self, "drop_prob", 0.05), insert_prob=getattr(self, "insert_prob", 0.0),
            smooth_low=getattr(self, "smooth_low", 0.0), smooth_high=getattr(self, "smooth_high", 0.1))

    def _merge_input_ids_with_audio_features(self, audio_features, num_audio_tokens, inputs_embeds, input_ids,
                                             attention_mask, labels):
        return _bridge.merge_input_ids_with_audio_features(
            audio_features, num_audio_tokens, inputs_embeds, input_ids, attention_mask, labels,
            self.tokenizer.default_speech_token, self.tokenizer.pad_token_id, self.tokenizer.default_ignore_token)

    # ------------------------------------------------------------------ shared front end
    def _encode(self, input_features, input_feature_length):
        """SenseVoice front end + encoder (ps-slm.py:427-447) — upstream of the bridge, unchanged."""
        speech = input_features
        B = speech.size(0)
        q = lambda ids: self.encoder.embed(torch.tensor([ids], device=speech.device)).repeat(B, 1, 1)  # noqa: E731
        speech = torch.cat([q([0]), q([1, 2]), q([2]), speech], dim=1)
        out, out_lens = self.encoder.encoder(speech, input_feature_length + 4)
        if isinstance(out, tuple):
            out = out[0]
        return out, out_lens

    def _text_only(self) -> bool:
        """The branch that never reads the encoder output: simulated posteriors from the transcript.  ``voca_trans``
        takes precedence over ``gt_emb`` in the reference's dispatch (ps-slm.py:456-458 / :587-589)."""
        return bool(self.ctc_posterior and self.gt_emb and not self.voca_trans)

    def _fused_bridge(self):
        if self._bridge is None:
            ctc_lo = self.encoder.ctc.ctc_lo
            self._bridge = _bridge.TasuBridge(
                ctc_lo.weight, ctc_lo.bias, self.encoder_projector, self.llm.get_input_embeddings().weight,
                self.tokenizer.default_speech_token, self.tokenizer.pad_token_id, self.tokenizer.default_ignore_token,
                blank_id=self.encoder.blank_id)
        return self._bridge

    def _bridge_outputs(self, raw_encoder_out, raw_encoder_out_lens, input_ids, attention_mask, labels, texts, noisy):
        """Steps 1–4 of the hot path with the reference's flag dispatch (ps-slm.py:456-528 / :587-658)."""
        table = self.llm.get_input_embeddings().weight
        fused_ok = (self.ctc_posterior and not self.voca_trans and not self.gt_emb and self.do_psd
                    and type(self.encoder_projector).__name__ == "EncoderProjectorLinearSiLU"
                    and not (torch.is_grad_enabled() and (table.requires_grad
                                                          or any(p.requires_grad for p in self.encoder_projector.parameters())))
                    and table.dtype in (torch.float32, torch.bfloat16))
        if fused_ok:                                           # shipped inference configuration
            emb, mask, out_labels, pos, _ = self._fused_bridge()(raw_encoder_out, raw_encoder_out_lens, input_ids,
                                                                 attention_mask, labels)
            return emb, mask, out_labels, pos
        blank = self.encoder.blank_id
        train_fused_ok = (self.ctc_posterior and not self.voca_trans and not self.gt_emb and self.do_psd
                          and type(self.encoder_projector).__name__ == "EncoderProjectorLinearSiLU"
                          and torch.is_grad_enabled() and any(p.requires_grad for p in self.encoder_projector.parameters())
                          and not raw_encoder_out.requires_grad and table.dtype in (torch.float32, torch.bfloat16))
        if train_fused_ok:
            # training on audio (ps-slm.py:450-454, :469-473, :482, :525-528 — the "half_audio_finetuned" recipe): the
            # no-grad head (fused ctc_lo statistics → exact decisions → collapse → kept-frame softmax GEMM → pooling)
            # never materialises the [B, T, 25055] posterior; the pooled rows enter the differentiable projector
            from .autograd import linear_silu_train_rows
            br = self._fused_bridge()
            pooled, mean, rstd, feat_len, max_len = br.compress_pooled(raw_encoder_out, raw_encoder_out_lens)
            pending = _bridge.begin_splice_plan(input_ids, attention_mask, feat_len, self.tokenizer.default_speech_token)
            if pooled.shape[0]:
                audio = linear_silu_train_rows(self.encoder_projector, pooled, mean, rstd, pooled.shape[0], table.dtype)
            else:
                audio = torch.zeros(0, table.shape[1], dtype=table.dtype, device=table.device)
            emb, mask, out_labels, pos, _ = _bridge.merge_packed_audio_rows(
                audio, feat_len, max_len, table, 1, input_ids, attention_mask, labels,
                self.tokenizer.default_speech_token, self.tokenizer.pad_token_id, self.tokenizer.default_ignore_token,
                pending=pending)
            return emb, mask, out_labels, pos
        rows_ok = (self._text_only() and self.token_row_path
                   and type(self.encoder_projector).__name__ == "EncoderProjectorLinearSiLU"
                   and table.dtype in (torch.float32, torch.bfloat16))
        if rows_ok:
            # text-only branch (ps-slm.py:459-468 / :589-598): the simulated posterior is one-hot plus a constant per
            # row, so it is handed to the projector as (token, hot, base) descriptors and never materialised
            ids_list = [self.encoder_tokenizer.encode(t) for t in texts]
            V = self.encoder_tokenizer.vocab_size
            hp = dict(drop_prob=getattr(self, "drop_prob", 0.05), smooth_low=getattr(self, "smooth_low", 0.0),
                      smooth_high=getattr(self, "smooth_high", 0.1))
            insert_prob = getattr(self, "insert_prob", 0.0)
            if noisy and int(max((len(i) for i in ids_list), default=0) * insert_prob) == 0:
                rows = _ops.sim_token_rows(_ops.TokenBatch(ids_list), V, input_ids.device, **hp)   # one native host call
            else:
                desc = (_sim.draw_noise_descriptors(ids_list, V, blank, insert_prob=insert_prob, **hp) if noisy
                        else _sim.clean_descriptors(ids_list))
                rows = _ops.group_token_rows(*desc, V, input_ids.device)
            feat_len = rows.lens // self.encoder_projector.k
            pending = _bridge.begin_splice_plan(input_ids, attention_mask, feat_len, self.tokenizer.default_speech_token)
            audio = self.encoder_projector.forward_token_rows(rows, out_dtype=table.dtype)
            emb, mask, out_labels, pos, _ = _bridge.merge_packed_audio_rows(
                audio, feat_len, max(rows.lens_host, default=0), table, 1, input_ids,
                attention_mask, labels, self.tokenizer.default_speech_token, self.tokenizer.pad_token_id,
                self.tokenizer.default_ignore_token, pending=pending)
            return emb, mask, out_labels, pos
        if raw_encoder_out is not None:                        # None on the text-only branch (encoder skipped)
            encoder_out = raw_encoder_out[:, 4:, :]
            encoder_out_lens = torch.clamp(raw_encoder_out_lens - 4, min=0)
        if self.ctc_posterior and self.voca_trans:                 # ps-slm.py:485-513 / :615-643
            tb = self._table_cache.get([table], lambda: (
                table.detach().contiguous() if table.dtype == torch.bfloat16
                else _ops.cast_rows(table.detach().contiguous(), torch.bfloat16)[0]),
                fresh=torch.is_grad_enabled() and table.requires_grad, verify=True)
            projector_outs, feat_len = _bridge.voca_trans_project(self.encoder_projector, encoder_out, encoder_out_lens, tb,
                                                                  self.do_psd, self.top1_emb)
            inputs_embeds = self.llm.get_input_embeddings()(input_ids)
            emb, mask, out_labels, pos, _ = self._merge_input_ids_with_audio_features(
                projector_outs, feat_len, inputs_embeds, input_ids, attention_mask, labels)
            return emb, mask, out_labels, pos
        if self.ctc_posterior:
            if self.gt_emb:                                        # voca_trans was handled above: this is _text_only()
                post, lens = (self.ctc_pseudo_posterior_noise(texts) if noisy else self.ctc_pseudo_posterior(texts))
                encoder_outs, feat_len = post.to(input_ids.device), lens.to(input_ids.device)
            else:
                logits = self.encoder.ctc.ctc_lo(raw_encoder_out)
                post = torch.softmax(logits, dim=-1)[:, 4:, :]
                if self.do_psd:
                    encoder_outs, feat_len = self.psd(post, encoder_out_lens, post, blank)
                else:
                    encoder_outs, feat_len = post, encoder_out_lens
        else:
            if self.do_psd:
                # raw-feature branch: segmentation from the fused CTC-head statistics, no [B, T, V] posterior
                ctc_lo = self.encoder.ctc.ctc_lo
                w_bf16, b_f32 = self._ctc_cache.get([ctc_lo.weight, ctc_lo.bias], lambda: (
                    _bridge.cast_weight_bf16(ctc_lo.weight),
                    ctc_lo.bias.detach().float().contiguous() if ctc_lo.bias is not None
                    else torch.zeros(ctc_lo.weight.shape[0], dtype=torch.float32, device=ctc_lo.weight.device)), verify=True)
                encoder_outs, feat_len = _bridge.psd_from_encoder(raw_encoder_out, raw_encoder_out_lens, w_bf16, b_f32, blank)
            else:
                encoder_outs, feat_len = encoder_out, encoder_out_lens
        if self.cross_attn and self.ctc_posterior:                 # ps-slm.py:475-480 / :605-610
            projector_outs = self.encoder_projector(encoder_outs, table.detach())
        else:
            projector_outs = self.encoder_projector(encoder_outs)
            feat_len = feat_len // self.encoder_projector.k
        inputs_embeds = self.llm.get_input_embeddings()(input_ids)
        emb, mask, out_labels, pos, _ = self._merge_input_ids_with_audio_features(
            projector_outs, feat_len, inputs_embeds, input_ids, attention_mask, labels)
        return emb, mask, out_labels, pos

    # ------------------------------------------------------------------ reference entry points
    def forward(self, input_ids: torch.LongTensor = None, input_features: Optional[torch.Tensor] = None,
                attention_mask: Optional[torch.Tensor] = None, input_feature_length: Optional[torch.Tensor] = None,
                position_ids=None, past_key_values=None, inputs_embeds=None, GT: Optional[List[str]] = None,
                labels: Optional[torch.LongTensor] = None, use_cache=None, output_attentions=None,
                output_hidden_states=None, return_dict=None):
        if self._text_only():
            # text-only branch: the encoder output is never used (ps-slm.py:459-468); skipping the dead
            # encoder pass changes no result
            raw, raw_lens = None, None
        else:
            raw, raw_lens = self._encode(input_features, input_feature_length)
        inputs_embeds, attention_mask, labels, position_ids = self._bridge_outputs(
            raw, raw_lens, input_ids, attention_mask, labels, GT, self.gt_emb_noise)
        model_outputs = self.llm(inputs_embeds=inputs_embeds, attention_mask=attention_mask, labels=labels,
                                 position_ids=position_ids)
        acc = -1
        if self.metric:
            with torch.no_grad():
                preds = torch.argmax(model_outputs.logits, -1)
                tgt = labels.detach()[:, 1:]
                keep = tgt != self.tokenizer.default_ignore_token
                acc = ((preds.detach()[:, :-1][keep] == tgt[keep]).sum().float() / keep.sum().float())
        return model_outputs, acc

    @torch.no_grad()
    def generate(self, input_ids: torch.LongTensor = None, input_features: Optional[torch.Tensor] = None,
                 attention_mask: Optional[torch.Tensor] = None, input_feature_length: Optional[torch.Tensor] = None,
                 position_ids=None, past_key_values=None, inputs_embeds=None, labels=None, use_cache=None,
                 output_attentions=None, output_hidden_states=None, return_dict=None, targets=None, **kwargs):
        texts = None
        if self._text_only():
            texts = [re.sub(r"[^A-Za-z\s.,!?]+", "", t).lower().strip() for t in targets]   # ps-slm.py:592-594
            raw, raw_lens = None, None
        else:
            raw, raw_lens = self._encode(input_features, input_feature_length)
        inputs_embeds, attention_mask, labels, position_ids = self._bridge_outputs(
            raw, raw_lens, input_ids, attention_mask, labels, texts, False)
        return self.llm.generate(
            inputs_embeds=inputs_embeds,
            max_new_tokens=kwargs.get("max_new_tokens", 200), num_beams=kwargs.get("num_beams", 4),
            do_sample=kwargs.get("do_sample", False), min_length=kwargs.get("min_length", 1),
            top_p=kwargs.get("top_p", 1.0), repetition_penalty=kwargs.get("repetition_penalty", 1.0),
            length_penalty=kwargs.get("length_penalty", 1.0), temperature=kwargs.get("temperature", 1.0),
            attention_mask=attention_mask, bos_token_id=self.tokenizer.bos_token_id,
            eos_token_id=self.tokenizer.eos_token_id, pad_token_id=self.tokenizer.pad_token_id)
