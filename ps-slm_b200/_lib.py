"""ctypes binding of libtasu_bridge.so (the C ABI declared in include/tasu_bridge.h).

The product path has NO fallback: if the shared library is missing or a call fails the
caller gets an exception, never a silent PyTorch/CPU substitute.
"""
import ctypes
import os
import subprocess
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p, POINTER

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libtasu_bridge.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

TASU_OK = 0
ABI_VERSION = 2
F32, BF16 = 0, 1
INPUT_PROBS, INPUT_LOGITS = 0, 1
EPI_NONE, EPI_BIAS, EPI_BIAS_SILU, EPI_BIAS_RELU, EPI_LNFOLD_SILU, EPI_LNFOLD, EPI_SOFTMAX = 0, 1, 2, 3, 4, 5, 6
SH_SPLICED_LEN, SH_LEFT_PADDING, SH_ERR_BOTH_SIDES, SH_TOTAL_SLOTS, SH_TOTAL_AUDIO, SH_N_SPEECH, SH_WORDS = 0, 1, 2, 3, 4, 5, 8
CH_N_OUT, CH_MAX_LEN, CH_IS_LOGPROB, CH_KEPT_FRAMES, CH_WORDS = 0, 1, 2, 3, 4
OPT_GEMM_PAIR, OPT_COUNT = 0, 1     # run-time options (tasu_set_option)
# layout words of the grouped kept-frame layout (tasu_group_plan)
GL_A2, GL_A4, GL_AX, GL_A_ROWS, GL_O4, GL_OX, GL_N2, GL_N4, GL_NS, GL_NXE, GL_WORDS = 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 16

# name -> (restype, argtypes); mirrors include/tasu_bridge.h one to one
_P, _I, _L, _F = c_void_p, c_int, c_int64, c_float
SIGNATURES = {
    "tasu_abi_version": (_I, []),
    "tasu_last_error": (c_char_p, []),
    "tasu_device_info": (_I, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "tasu_set_option": (_I, [_I, _I]),
    "tasu_get_option": (_I, [_I]),
    "tasu_frame_stats": (_I, [_P, _I, _I, _I, _I, _I, _L, _L, _I, _P, _P, _P, _P, _P, _P, _P]),
    "tasu_collapse_plan": (_I, [_P, _P, _P, _P, _P, _I, _P, _I, _I, _I, _F, _P, _P, _P, _P, _P, _P, _P]),
    "tasu_collapse_plan_scan": (_I, [_P, _P, _P, _P, _P, _I, _P, _I, _I, _I, _F, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "tasu_row_norm_max": (_I, [_P, _I, _I, _I, _L, _P, _P]),
    "tasu_flag_ambiguous_frames": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _L, _I, _P, _P, _F, _I, _I, _I, _I, _F, _P, _P, _P, _P, _P, _P]),
    "tasu_cast_rows_sumsq": (_I, [_P, _L, _I, _L, _P, _L, _P, _P]),
    "tasu_ctc_head_refine_workspace": (_L, [_L]),
    "tasu_ctc_head_refine": (_I, [_P, _I, _L, _P, _L, _P, _I, _I, _I, _P, _P, _P, _L, _P, _P, _P, _P, _P, _L, _P]),
    "tasu_collapse_scan": (_I, [_P, _P, _P, _I, _P, _P, _P, _P, _P]),
    "tasu_gather_kept_rows": (_I, [_P, _L, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _L, _L, _P, _L, _P, _P,
                                   _P, _P, _P, _P, _P, _P, _F, _P]),
    "tasu_kept_frame_index": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _L, _L, _P, _P, _P]),
    "tasu_pool_tail": (_I, [_P, _L, _I, _L, _L, _P, _P, _P, _P, _P, _P, _F, _P]),
    "tasu_group_plan": (_I, [_P, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "tasu_gather_kept_rows_grouped": (_I, [_P, _L, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _L, _L, _L,
                                           _P, _L, _P, _P, _P, _P, _P, _P, _P, _P, _P, _F, _P]),
    "tasu_gemm_softmax_grouped_parts": (_I, [_I]),
    "tasu_gemm_softmax_grouped": (_I, [_P, _L, _P, _L, _P, _L, _I, _I, _I, _I, _P, _P, _P, _P, _P, _L, _P]),
    "tasu_group_ln_finish": (_I, [_P, _L, _I, _P, _I, _L, _P, _P, _F, _P]),
    "tasu_segment_meanpool": (_I, [_P, _I, _I, _I, _I, _L, _L, _P, _P, _P, _P, _P, _P, _I, _L, _L, _P, _I, _L, _P, _P, _F, _P]),
    "tasu_sim_posterior_rows": (_I, [_P, _P, _P, _P, _L, _I, _P, _I, _L, _P, _P, _F, _P]),
    "tasu_fingerprint": (_I, [_P, _P, _I, _P, _P]),
    "tasu_cast_rows": (_I, [_P, _I, _L, _I, _L, _P, _I, _L, _P, _P, _F, _P]),
    "tasu_softmax_rows": (_I, [_P, _I, _L, _L, _I, _P, _P, _P, _L, _P]),
    "tasu_fold_layernorm": (_I, [_P, _L, _P, _P, _P, _I, _I, _P, _L, _P, _P, _P]),
    "tasu_gemm_bf16_tn": (_I, [_P, _L, _P, _L, _P, _I, _L, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "tasu_gemm_streamk_workspace": (_L, []),
    "tasu_gemm_bf16_tn_streamk": (_I, [_P, _L, _P, _L, _P, _I, _L, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _L, _P]),
    "tasu_gemm_streamk_schedule_host": (_I, [_I, _I, _I, _I, _P, _P]),
    "tasu_attn_softmax_pv": (_I, [_P, _L, _P, _L, _I, _I, _I, _I, _P, _P, _L, _P, _L, _P]),
    "tasu_attn_split_plan": (_L, [_I, _I, _I, _I, POINTER(c_int)]),
    "tasu_attn_softmax_pv_ws": (_I, [_P, _L, _P, _L, _I, _I, _I, _I, _P, _P, _L, _P, _L, _P, _L, _P]),
    "tasu_gemm_bf16_f32": (_I, [_P, _L, _I, _P, _L, _I, _P, _L, _I, _I, _I, _P]),
    "tasu_ctc_head_stats_workspace": (_L, [_I, _I, _I]),
    "tasu_ctc_head_stats": (_I, [_P, _L, _P, _L, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _L, _P]),
    "tasu_gemm_bf16_tn_simt": (_I, [_P, _L, _P, _L, _P, _I, _L, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "tasu_attn_score_grad": (_I, [_P, _L, _P, _L, _P, _P, _L, _I, _L, _I, _P, _L, _P]),
    "tasu_transpose_cast": (_I, [_P, _I, _L, _L, _L, _P, _P, _L, _P]),
    "tasu_silu_fwd": (_I, [_P, _L, _P, _P]),
    "tasu_silu_bwd": (_I, [_P, _P, _L, _I, _P, _P, _P, _L, _P, _P, _P]),
    "tasu_colsum": (_I, [_P, _I, _L, _I, _L, _P, _P]),
    "tasu_linear_silu_wgrad_finish": (_I, [_P, _L, _P, _L, _P, _P, _P, _P, _I, _I, _P, _L, _P, _P, _P]),
    "tasu_split_bf16x3": (_I, [_P, _I, _L, _I, _L, _P, _I, _P, _L, _P, _P, _F, _P, _P]),
    "tasu_sum_epilogue": (_I, [_P, _I, _L, _I, _I, _L, _I, _P, _P, _P, _P, _P, _I, _L, _P]),
    "tasu_host_group_tokens": (_I, [_P, _L, _I, _P, _P, _P, _P]),
    "tasu_host_sim_token_rows": (_I, [_P, _P, _P, _I, _I, _F, _F, _F, _P, _L, _P, _P]),
    "tasu_linear_rowdots": (_I, [_P, _L, _P, _P, _P, _I, _I, _P, _P, _P]),
    "tasu_tokrow_fwd": (_I, [_P, _L, _P, _P, _P, _P, _P, _P, _P, _P, _I, _L, _I, _I, _F, _P, _P, _P, _P, _P, _P]),
    "tasu_tokrow_cols_workspace": (_L, [_I, _I]),
    "tasu_tokrow_rows_fwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _L, _I, _I, _F, _P, _P, _P, _P, _P, _P]),
    "tasu_tokrow_cols": (_I, [_P, _L, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _L, _P]),
    "tasu_tokrow_bwd_workspace": (_L, [_I, _I]),
    "tasu_tokrow_bwd_rows": (_I, [_P, _P, _L, _I, _P, _P, _P, _P, _I, _P, _P, _P, _P, _L, _P]),
    "tasu_tokrow_wgrad_finish": (_I, [_P, _P, _I, _P, _P, _L, _P, _P, _P, _P, _I, _I, _P, _L, _P, _P, _P]),
    "tasu_tokrow_train_workspace": (_L, [_L, _I, _I, _I, _I]),
    "tasu_tokrow_linear_silu_fwd": (_I, [_P, _L, _P, _P, _P, _P, _L, _P, _P, _P, _P, _P, _P, _I, _L, _I, _I, _I, _F,
                                         _P, _P, _P, _P, _P, _I, _L, _P, _L, _P]),
    "tasu_tokrow_linear_silu_bwd": (_I, [_P, _I, _L, _P, _P, _P, _P, _P, _L, _P, _P, _P, _L, _P, _P, _P, _I, _L, _I, _I,
                                         _I, _P, _L, _P, _P, _P, _P, _L, _P, _I, _P, _L, _P]),
    "tasu_splice_rowstat": (_I, [_P, _P, _I, _I, _I, _L, _P, _P]),
    "tasu_splice_plan": (_I, [_P, _P, _I, _I, _I, _L, _P, _I, _L, _P, _P, _P, _P, _P]),
    "tasu_splice_header": (_I, [_P, _P, _I, _L, _I, _I, _P, _P, _P, _P]),
    "tasu_splice_plan_header": (_I, [_P, _P, _I, _I, _I, _L, _P, _I, _L, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "tasu_splice_scatter": (_I, [_P, _P, _I, _P, _I, _I, _I, _I, _L, _P, _I, _L, _P, _I, _L, _L, _I, _I,
                                 _P, _P, _P, _P, _P, _P, _I, _L, _L, _P, _P, _P, _P, _P, _P, _P, _P]),
    "tasu_splice_scatter_perm": (_I, [_P, _P, _I, _P, _I, _I, _I, _I, _L, _P, _I, _L, _P, _P, _I, _L, _L, _I, _I,
                                      _P, _P, _P, _P, _P, _P, _I, _L, _L, _P, _P, _P, _P, _P, _P, _P, _P]),
    "tasu_flat_scale_cast": (_I, [_P, _I, _P, _I, _L, _F, _P]),
    "tasu_packed_select": (_I, [_P, _I, _L, _L, _I, _P, _I, _I, _L, _P, _I, _P, _L, _L, _P, _P, _P, _P]),
    "tasu_splice_text_grad": (_I, [_P, _I, _L, _P, _L, _I, _P, _L, _L, _P]),
    "tasu_gather_rows": (_I, [_P, _I, _L, _P, _L, _I, _P, _L, _P]),
}

_lib = None


class TasuError(RuntimeError):
    pass


def build(verbose=False):
    """Compile libtasu_bridge.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-j8", "-C", CSRC_DIR], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise TasuError("building libtasu_bridge.so failed")
    return LIB_PATH


def lib():
    """Load the shared library (once). Raises TasuError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise TasuError(
            "libtasu_bridge.so not found at %s — run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU/PyTorch fallback for the bridge path)" % LIB_PATH)
    handle = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)          # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    if handle.tasu_abi_version() != ABI_VERSION:
        raise TasuError("libtasu_bridge.so ABI version mismatch")
    _lib = handle
    return _lib


def check(rc, what):
    if rc != TASU_OK:
        msg = lib().tasu_last_error()
        raise TasuError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else ""))
