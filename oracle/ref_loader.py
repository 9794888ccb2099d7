"""Load the UNMODIFIED reference (PigeonDan1/ps-slm) for oracle pinning.

Test infrastructure only.  Uses ``/root/reference`` where it exists (the build
container), else the byte-for-byte copies ``oracle/vendor_ref.py`` staged under
the git-ignored ``oracle/_ref/`` (they travel to the GPU box with the working
tree); ``available()`` is False when neither exists and every caller must skip.  Recipe: put ``Multitask/`` on sys.path, stub the two
absent third-party modules that are imported at module scope but never used on
the bridge path (``peft``: Multitask/model/ps-slm.py:16,
Multitask/utils/config_utils.py:9-13; ``omegaconf``: utils/config_utils.py:15),
and load ``model/ps-slm.py`` by file path (hyphen in the name, same trick as
Multitask/utils/dataset_utils.py:14-25).  The bridge methods are then called
unbound on a light fake ``self``.
"""
import importlib.machinery
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "Multitask")   # oracle/vendor_ref.py
REF_ROOT = os.environ.get("TASU_REFERENCE_ROOT") or (
    "/root/reference/Multitask" if os.path.isfile("/root/reference/Multitask/model/ps-slm.py") else _STAGED)

_cache = {}


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "model", "ps-slm.py"))


def _stub(name, attrs):
    if name in sys.modules:
        return
    m = types.ModuleType(name)
    for a in attrs:
        setattr(m, a, type(a, (), {}))
    sys.modules[name] = m


def load():
    """Return (ps_slm_module, projector_module) of the reference."""
    if "mods" in _cache:
        return _cache["mods"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _stub("peft", ["PeftModel", "LoraConfig", "TaskType", "get_peft_model",
                   "prepare_model_for_kbit_training", "AdaptionPromptConfig",
                   "PrefixTuningConfig"])
    _stub("omegaconf", ["OmegaConf"])
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    loader = importlib.machinery.SourceFileLoader(
        "tasu_reference_ps_slm", os.path.join(REF_ROOT, "model", "ps-slm.py"))
    spec = importlib.util.spec_from_loader(loader.name, loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    ploader = importlib.machinery.SourceFileLoader(
        "tasu_reference_projector", os.path.join(REF_ROOT, "model", "projector.py"))
    pspec = importlib.util.spec_from_loader(ploader.name, ploader)
    pmod = importlib.util.module_from_spec(pspec)
    ploader.exec_module(pmod)
    _cache["mods"] = (mod, pmod)
    return mod, pmod


class FakeSenseVoiceTokenizer:
    """Stand-in for Multitask/model/tokenizer.py:5-29 (BPE asset is not in the
    repo): a "text" is a whitespace-separated list of integer token ids."""

    def __init__(self, vocab_size=25055):
        self._v = vocab_size

    def encode(self, text):
        return [int(t) for t in text.split()]

    @property
    def vocab_size(self):
        return self._v


class FakeLLMTokenizer:
    def __init__(self, speech_id, pad_id, ignore=-100):
        self.default_speech_token = speech_id
        self.pad_token_id = pad_id
        self.default_ignore_token = ignore


class FakeSelf(nn.Module):
    """Just enough of ``slam_model_asr`` for the unbound bridge methods."""

    def __init__(self, vocab_size=25055, blank_id=0, speech_id=151665, pad_id=151643):
        super().__init__()
        self._dummy = nn.Parameter(torch.zeros(1))
        self.encoder_tokenizer = FakeSenseVoiceTokenizer(vocab_size)
        self.encoder = types.SimpleNamespace(blank_id=blank_id)
        self.tokenizer = FakeLLMTokenizer(speech_id, pad_id)


def ref_psd(encoder_out, lens, posterior, blank_id=0, blank_threshold=0.90):
    """Multitask/model/ps-slm.py:237-317, unmodified (stdout prints swallowed)."""
    mod, _ = load()
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return mod.slam_model_asr.psd(FakeSelf(), encoder_out, lens, posterior, blank_id, blank_threshold)


def ref_sim_clean(texts, vocab_size=25055):
    """Multitask/model/ps-slm.py:337-358, unmodified."""
    mod, _ = load()
    return mod.slam_model_asr.ctc_pseudo_posterior(FakeSelf(vocab_size), texts)


def ref_sim_noise(texts, vocab_size=25055, blank_id=0, **attrs):
    """Multitask/model/ps-slm.py:360-409, unmodified. ``attrs`` = drop_prob,
    insert_prob, smooth_low, smooth_high overrides (read with getattr at :372-375)."""
    mod, _ = load()
    fake = FakeSelf(vocab_size, blank_id)
    for k, v in attrs.items():
        setattr(fake, k, v)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return mod.slam_model_asr.ctc_pseudo_posterior_noise(fake, texts)


def ref_merge(audio_features, num_audio_tokens, inputs_embeds, input_ids, attention_mask, labels,
              speech_id=151665, pad_id=151643):
    """Multitask/model/ps-slm.py:679-873, unmodified."""
    mod, _ = load()
    fake = FakeSelf(speech_id=speech_id, pad_id=pad_id)
    return mod.slam_model_asr._merge_input_ids_with_audio_features(
        fake, audio_features, num_audio_tokens, inputs_embeds, input_ids, attention_mask, labels)


def ref_projector(kind, encoder_dim, llm_dim, ds_rate=1):
    """Construct a reference projector (Multitask/model/projector.py)."""
    _, pmod = load()
    cfg = types.SimpleNamespace(encoder_dim=encoder_dim, llm_dim=llm_dim,
                                encoder_projector_ds_rate=ds_rate)
    cls = {"linear-silu": pmod.EncoderProjectorLinearSiLU,
           "linear": pmod.EncoderProjectorConcat,
           "simple_linear": pmod.EncoderProjectorLinear,
           "cross-attention": pmod.EncoderProjectorCTCCA}[kind]
    return cls(cfg)


def load_dataset_module():
    """The reference's dataset/speech_dataset_large.py with its unused heavy imports (whisper, kaldiio, torchaudio)
    stubbed — only ``MultiTaskDataset.collator`` / ``.pad``, ``MultiTaskDynamicBatchDataset`` and ``window_class``
    (Multitask/dataset/speech_dataset_large.py:190-338) are exercised."""
    if "dataset" in _cache:
        return _cache["dataset"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    added = []
    for name in ("whisper", "kaldiio", "torchaudio", "torchaudio.compliance", "torchaudio.compliance.kaldi"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []                                   # lets "import a.b.c" resolve through sys.modules
            m.__spec__ = importlib.machinery.ModuleSpec(name, None)
            sys.modules[name] = m
            added.append(name)
    if "torchaudio" in added:
        sys.modules["torchaudio"].compliance = sys.modules["torchaudio.compliance"]
        sys.modules["torchaudio.compliance"].kaldi = sys.modules["torchaudio.compliance.kaldi"]
    loader = importlib.machinery.SourceFileLoader(
        "tasu_reference_dataset", os.path.join(REF_ROOT, "dataset", "speech_dataset_large.py"))
    spec = importlib.util.spec_from_loader(loader.name, loader)
    mod = importlib.util.module_from_spec(spec)
    try:
        loader.exec_module(mod)
    finally:
        for name in added:                                    # the stubs must not leak: transformers probes
            sys.modules.pop(name, None)                       # importlib.util.find_spec("torchaudio") later on
    _cache["dataset"] = mod
    return mod


def load_voca_fixed():
    """The reference's model/ps-slm.py with ONE line added in memory: its ``voca_trans`` branch reads ``encoder_outs`` /
    ``encoder_feature_length`` before assigning them (ps-slm.py:488, :618 — UnboundLocalError as shipped); the fix binds
    them to ``encoder_out`` / ``encoder_out_lens`` right after the branch's print statement.  Nothing else differs from
    the file on disk, which is never modified."""
    if "voca" in _cache:
        return _cache["voca"]
    load()                                                     # stubs + sys.path
    path = os.path.join(REF_ROOT, "model", "ps-slm.py")
    src = open(path).read()
    marker = 'print("Vocabulary Transform is ready ...")'
    assert src.count(marker) == 2, "reference layout changed"
    src = src.replace(marker, marker + "; encoder_outs, encoder_feature_length = encoder_out, encoder_out_lens")
    mod = types.ModuleType("tasu_reference_ps_slm_voca_fixed")
    mod.__file__ = path
    exec(compile(src, path, "exec"), mod.__dict__)
    _cache["voca"] = mod
    return mod
