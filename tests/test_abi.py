"""CPU: the C-ABI library loads, exports every symbol include/tasu_bridge.h declares with the
declared arity, validates arguments without touching a GPU, and the host layer refuses CPU
tensors instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "tasu_bridge.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"(?:int64_t|int|const char\*)\s+(tasu_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else args.count(",") + 1
        out[m.group(1)] = n
    return out


@pytest.fixture(scope="module")
def lib():
    import ps_slm_b200
    if not os.path.isfile(ps_slm_b200._lib.LIB_PATH):
        ps_slm_b200.build()
    return ps_slm_b200.lib()


def test_header_symbols_exported(lib):
    import ps_slm_b200._lib as L
    funcs = _header_functions()
    assert len(funcs) >= 17
    raw = ctypes.CDLL(L.LIB_PATH)
    for name, nargs in funcs.items():
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
        assert name in L.SIGNATURES, f"{name} has no ctypes signature"
        assert len(L.SIGNATURES[name][1]) == nargs, f"{name}: header has {nargs} args, binding {len(L.SIGNATURES[name][1])}"
    assert set(L.SIGNATURES) == set(funcs)
    assert lib.tasu_abi_version() == 2


def test_argument_validation_without_gpu(lib):
    import ps_slm_b200._lib as L
    # V = 0
    rc = lib.tasu_frame_stats(None, L.F32, L.INPUT_PROBS, 1, 1, 0, 0, 0, 0, None, None, None, None, None, None, None)
    assert rc == -1 and b"tasu_frame_stats" in lib.tasu_last_error()
    # unaligned GEMM pitch
    rc = lib.tasu_gemm_bf16_tn(16, 3, 16, 8, 16, L.F32, 8, 4, 4, 8, 0, None, None, None, None, None, None)
    assert rc == -1
    # LN-fold epilogue without its vectors
    rc = lib.tasu_gemm_bf16_tn(16, 8, 16, 8, 16, L.F32, 8, 4, 4, 8, L.EPI_LNFOLD_SILU, 16, None, None, None, None, None)
    assert rc == -1 and b"LN-fold" in lib.tasu_last_error()
    rc = lib.tasu_segment_meanpool(None, 7, 1, 1, 1, 1, 1, None, None, None, None, None, None, 0, 1, 1, None, 0, 1, None, None, 1e-5, None)
    assert rc == -1
    rc = lib.tasu_splice_plan(None, None, 3, 1, 1, 0, None, 0, 1, None, None, None, None, None)
    assert rc == -1
    with pytest.raises(L.TasuError):
        L.check(rc, "tasu_splice_plan")
    # empty problems are fine without a device
    assert lib.tasu_collapse_plan(None, None, None, None, None, 0, None, 0, 10, 0, 0.9, None, None, None, None, None, None, None) == 0
    assert lib.tasu_cast_rows(None, 0, 0, 8, 8, None, 1, 8, None, None, 1e-5, None) == 0


def test_no_cpu_fallback():
    import ps_slm_b200.bridge as bridge
    import ps_slm_b200._lib as L
    x = torch.rand(1, 4, 8)
    with pytest.raises(L.TasuError):
        bridge.psd(x, torch.tensor([4]), x)
    with pytest.raises(L.TasuError):
        bridge.merge_input_ids_with_audio_features(torch.zeros(1, 2, 4), torch.tensor([2]), torch.zeros(1, 3, 4),
                                                   torch.tensor([[1, 9, 2]]), torch.ones(1, 3, dtype=torch.bool), None, 9, 0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "ps-slm_b200")
    for name in os.listdir(pkg):
        if name.endswith(".py"):
            src = open(os.path.join(pkg, name)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), name


def test_projector_state_dict_names():
    import types
    import ps_slm_b200.projector as P
    cfg = types.SimpleNamespace(encoder_dim=25055, llm_dim=1536, encoder_projector_ds_rate=2)
    m = P.EncoderProjectorLinearSiLU(cfg)
    sd = m.state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == {
        "norm.weight": (25055,), "norm.bias": (25055,), "ffn.0.weight": (2048, 25055), "ffn.0.bias": (2048,),
        "ffn.2.weight": (1536, 2048), "ffn.2.bias": (1536,)}
    assert sum(p.numel() for p in m.parameters()) == 54512062 and m.k == 1
    assert bool((sd["ffn.2.bias"] == 0).all())
    c = P.EncoderProjectorConcat(types.SimpleNamespace(encoder_dim=512, llm_dim=1536, encoder_projector_ds_rate=2))
    assert set(c.state_dict()) == {"linear1.weight", "linear1.bias", "linear2.weight", "linear2.bias"} and c.k == 2
    assert tuple(c.linear1.weight.shape) == (2048, 1024)
    ca = P.EncoderProjectorCTCCA(types.SimpleNamespace(encoder_dim=25055, llm_dim=1536, encoder_projector_ds_rate=1))
    assert {k: tuple(v.shape) for k, v in ca.state_dict().items()} == {"W_q.weight": (1536, 25055)} and ca.n_heads == 8
    assert not hasattr(ca, "k")                       # like the reference: the cross-attention branch never reads .k
    s = P.EncoderProjectorLinear(types.SimpleNamespace(encoder_dim=512, llm_dim=151644, encoder_projector_ds_rate=1))
    assert set(s.state_dict()) == {"map.weight", "map.bias"} and tuple(s.map.weight.shape) == (151644, 512)


def test_options_defaults_and_validation(lib):
    """Run-time options: the CTA-pair GEMM is on by default (0 selects the one-CTA kernel); unknown ids are rejected."""
    import ps_slm_b200._lib as L
    import ps_slm_b200.ops as ops
    assert "TASU_GEMM_PAIR" not in os.environ
    assert ops.get_option(L.OPT_GEMM_PAIR) == 1
    ops.set_option(L.OPT_GEMM_PAIR, 0)
    assert ops.get_option(L.OPT_GEMM_PAIR) == 0
    ops.set_option(L.OPT_GEMM_PAIR, 1)
    assert ops.get_option(L.OPT_GEMM_PAIR) == 1
    assert lib.tasu_set_option(L.OPT_COUNT, 1) == -1 and b"unknown option" in lib.tasu_last_error()
    assert lib.tasu_get_option(-3) == -1
    with pytest.raises(L.TasuError):
        ops.set_option(99, 1)
