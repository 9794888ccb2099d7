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
    params = [torch.nn.Parameter(torch.zeros(3))]
    with pytest.raises(ValueError):
        D.enable_overlapped_allreduce(True)                                      # the parameters are required
    D.enable_overlapped_allreduce(True, params=params)
    try:
        assert D.overlap_hook() is None                                          # single process: nothing to overlap
        assert D.allreduce_gradients(params) == []
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


# ---------------------------------------------------------------------------------------------------------------
# EXPERIMENTAL stream-K GEMM: the schedule is plain integer logic shared by host and kernel — provable on the CPU
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("num_tiles,k_blocks,grid", [
    (528, 392, 148),      # projector GEMM-1 of the headline batch: 3 full waves + 84 tiles cut along K
    (444, 392, 148),      # exact multiple of the grid: no stream-K pieces at all
    (445, 392, 148),      # one leftover tile: cut 4 ways
    (591, 32, 148),       # 147 leftover tiles: every CTA gets almost a whole tile
    (7, 17, 148), (1, 1, 148), (3, 2, 148), (149, 1, 148), (37, 5, 8), (0, 9, 148), (1000, 33, 7),
])
def test_streamk_schedule_covers_every_k_block_once(num_tiles, k_blocks, grid):
    _check_streamk_schedule(num_tiles, k_blocks, grid)


def test_streamk_schedule_random_problems():
    rng = np.random.default_rng(7)
    for _ in range(150):
        grid = int(rng.choice([1, 2, 7, 16, 74, 132, 148]))
        _check_streamk_schedule(int(rng.integers(0, 6 * grid + 3)), int(rng.integers(1, 400)), grid)


def _check_streamk_schedule(num_tiles, k_blocks, grid):
    import ps_slm_b200.ops as ops
    FULL, CONTRIB, FINISH = 0, 1, 2
    dp_tiles, per_cta = ops.streamk_schedule(num_tiles, k_blocks, grid)
    floor_dp = (num_tiles // grid) * grid
    # a last wave that is at least 3/4 full is not cut
    assert dp_tiles == (num_tiles if 4 * (num_tiles - floor_dp) >= 3 * grid else floor_dp)
    cover = {}                       # tile -> list of (kb0, kb1, cta, kind, n_contrib)
    for cta, pieces in enumerate(per_cta):
        assert len(pieces) <= 2
        kinds = [p[3] for p in pieces]
        assert kinds.count(CONTRIB) <= 1, "one partial-accumulator slot per CTA"
        if CONTRIB in kinds:
            assert kinds[0] == CONTRIB, "a CTA contributes BEFORE it finishes (no wait chains)"
        for tile, kb0, kb1, kind, n_contrib in pieces:
            assert dp_tiles <= tile < num_tiles and 0 <= kb0 < kb1 <= k_blocks
            assert kind == (FULL if (kb0 == 0 and kb1 == k_blocks) else FINISH if kb1 == k_blocks else CONTRIB)
            cover.setdefault(tile, []).append((kb0, kb1, cta, kind, n_contrib))
    assert sorted(cover) == list(range(dp_tiles, num_tiles)), "every tile of the ragged wave is scheduled"
    for tile, ps in cover.items():
        ps.sort()
        assert ps[0][0] == 0 and ps[-1][1] == k_blocks
        for a, b in zip(ps, ps[1:]):
            assert a[1] == b[0], "pieces of a tile are contiguous and disjoint"
        ctas = [p[2] for p in ps]
        assert ctas == list(range(ctas[0], ctas[0] + len(ps))), "consecutive CTAs, ascending with K"
        assert len(ps) <= 6
        # exactly the last piece finishes (or the tile is whole) and it names the CTAs below it as its contributors
        assert [p[3] for p in ps[:-1]] == [CONTRIB] * (len(ps) - 1)
        assert ps[-1][3] == (FULL if len(ps) == 1 else FINISH)
        assert ps[-1][4] == len(ps) - 1
    # balance: no CTA gets more than one tile's worth of K-blocks in the tail
    for pieces in per_cta:
        assert sum(p[2] - p[1] for p in pieces) <= k_blocks


def test_pretrained_ctc_head_checkpoint_format(tmp_path):
    """setup_encoder_projector('simple_linear', ctc_linear=...) reads the reference's checkpoint layout (ps-slm.py:67-85):
    optionally wrapped in {"model": ...}, keys ctc_head.weight / ctc_head.bias, strict load, frozen with the encoder."""
    import ps_slm_b200.model as M
    w, b = torch.randn(11, 6), torch.randn(11)
    mc = types.SimpleNamespace(encoder_projector="simple_linear", encoder_dim=6, llm_dim=11, encoder_projector_ds_rate=1)
    for wrap in (True, False):
        state = {"ctc_head.weight": w, "ctc_head.bias": b, "unrelated": torch.zeros(1)}
        path = str(tmp_path / ("ckpt%d.pt" % wrap))
        torch.save({"model": state} if wrap else state, path)
        mc.ctc_linear = path
        for freeze in (True, False):
            proj = M.setup_encoder_projector(types.SimpleNamespace(freeze_encoder=freeze, freeze_projector=False), mc)
            assert torch.equal(proj.map.weight, w) and torch.equal(proj.map.bias, b)
            assert all(p.requires_grad != freeze for p in proj.parameters())
            assert proj.training != freeze
    torch.save({"ctc_lo.weight": w, "ctc_lo.bias": b}, str(tmp_path / "bad.pt"))
    mc.ctc_linear = str(tmp_path / "bad.pt")
    with pytest.raises(KeyError):
        M.setup_encoder_projector(types.SimpleNamespace(freeze_encoder=False, freeze_projector=False), mc)


def test_train_eval_switch_invalidates_weight_caches():
    """ProjectorCache: a train() / eval() switch of any bridge module drops every cached weight copy; a training forward
    (fresh=True) never caches — the copies of one evaluation phase cannot survive the optimizer steps after it."""
    import ps_slm_b200.bridge as bridge
    import ps_slm_b200.projector as P
    p = torch.nn.Parameter(torch.zeros(3))
    cache = bridge.ProjectorCache()
    built = []
    build = lambda: built.append(1) or len(built)      # noqa: E731
    assert cache.get([p], build) == 1 and cache.get([p], build) == 1
    with torch.no_grad():
        p.data.copy_(torch.ones(3))                    # invisible to torch's version counter (what ZeRO does)
    assert cache.get([p], build) == 1                  # ... so only an invalidation (or the fingerprint on CUDA) rebuilds
    proj = P.EncoderProjectorLinear(types.SimpleNamespace(encoder_dim=4, llm_dim=5, encoder_projector_ds_rate=1))
    proj.train()
    assert cache.get([p], build) == 2
    proj.eval()
    assert cache.get([p], build) == 3 and cache.get([p], build) == 3
    assert cache.get([p], build, fresh=True) == 4 and cache.get([p], build, fresh=True) == 5
    assert cache.get([p], build) == 6                  # a fresh build leaves nothing behind


def test_bind_rank_to_cpu_slice_partitions_the_allowed_cpus():
    """One process per GPU on one box: every rank gets its own contiguous, disjoint slice of the CPUs this process may
    use (ps-slm_b200/dist.py); nothing changes when there are fewer than two CPUs per rank or a single rank."""
    import os

    import ps_slm_b200.dist as D
    if not hasattr(os, "sched_getaffinity"):
        pytest.skip("no CPU affinity on this platform")
    before = sorted(os.sched_getaffinity(0))
    try:
        assert D.bind_rank_to_cpu_slice(0, 1) is None and sorted(os.sched_getaffinity(0)) == before
        world = max(2, min(8, len(before) // 2))
        if len(before) // world < 2:
            assert D.bind_rank_to_cpu_slice(0, world) is None
            return
        seen = []
        for r in range(world):
            os.sched_setaffinity(0, before)
            mine = D.bind_rank_to_cpu_slice(r, world)
            assert mine == sorted(os.sched_getaffinity(0)) and len(mine) == len(before) // world
            seen += mine
        assert len(set(seen)) == len(seen) and set(seen) <= set(before)
        os.sched_setaffinity(0, before)
        assert D.bind_rank_to_cpu_slice(0, 10 * len(before)) is None
    finally:
        os.sched_setaffinity(0, before)


def _grouped_layout_host(runs):
    """Host restatement of the layout words tasu_group_plan writes (include/tasu_bridge.h, step 2b') for a list of kept
    run lengths: (A rows, pooled rows incl. the per-frame rows of the long runs, pooled rows the projector reads)."""
    up = lambda v, a: (v + a - 1) // a * a
    ns = sum(1 for n in runs if n == 1 or n > 4)
    n2 = sum(1 for n in runs if n == 2)
    n4 = sum(1 for n in runs if 3 <= n <= 4)
    nxe = sum(n - 1 for n in runs if n > 4)
    a2 = up(ns, 128)
    a4 = up(a2 + 2 * n2, 128)
    ax = up(a4 + 4 * n4, 128)
    ox = a2 + up(n2, 64) + up(n4, 32)
    return ax + nxe, ox + nxe, ox


def test_grouped_layout_fits_the_capacities_for_any_run_length_mix():
    """ops.grouped_capacities / the projector's row capacity (N_out + 256) hold every batch whose kept-frame and candidate
    counts are within the capacities the host sized the buffers for — including the adversarial mixes (all 3-frame runs:
    one zero row each; all long runs: every extra frame keeps its own output row)."""
    import ps_slm_b200.ops as ops
    rng = np.random.default_rng(5)
    mixes = [[3] * 4000, [1] * 5000, [2] * 3001, [4] * 999, [9] * 700, [5, 1, 2, 3] * 500, [129] * 3, [1]]
    for _ in range(200):
        k = int(rng.integers(1, 3000))
        p = rng.dirichlet(np.ones(8))
        mixes.append(rng.choice([1, 2, 3, 4, 5, 6, 17, 200], size=k, p=p).tolist())
    for runs in mixes:
        n_out, n_frames = len(runs), sum(runs)
        a_rows, p_rows, proj_rows = _grouped_layout_host(runs)
        cap_a, cap_p = ops.grouped_capacities(n_frames, n_out)
        assert a_rows <= cap_a and p_rows <= cap_p and proj_rows <= n_out + 256
        assert proj_rows <= cap_p                          # the projector's rows lie inside the pooled matrix


def test_attention_key_split_plan():
    """tasu_attn_split_plan (HOST): the number of key ranges per (row tile, head) item of the cross-attention kernel —
    never more than there are key tiles or than 8, never a worse last wave than without a split, 6 at the config-2 size
    (616 items on 148 SMs: 4.16 -> 24.97 waves), workspace = items x splits x 128 x (dp + 2) floats."""
    import ctypes
    import math
    import ps_slm_b200._lib as L
    lib = L.lib()

    def plan(N, V2, h, d):
        s = ctypes.c_int(0)
        ws = int(lib.tasu_attn_split_plan(N, V2, h, d, ctypes.byref(s)))
        return s.value, ws

    assert plan(9856, 151936, 8, 192)[0] == 6
    rng = np.random.default_rng(11)
    sms = 148                                              # the library's answer without a device
    for _ in range(300):
        N, V2 = int(rng.integers(1, 20000)), int(rng.integers(1, 200000))
        h, d = int(rng.choice([1, 2, 4, 8, 12])), int(rng.choice([64, 128, 192, 256]))
        S, ws = plan(N, V2, h, d)
        items, J = (N + 127) // 128 * h, (V2 + 127) // 128
        assert 1 <= S <= min(8, J)
        assert ws == (items * S * 128 * (d + 2) * 4 if S > 1 else 0)
        time_of = lambda s: math.ceil(items * s / sms) / s   # waves x item length, in units of an unsplit item
        assert time_of(S) <= time_of(1) + 1e-12
    assert plan(0, 1000, 8, 192) == (1, 0)


@pytest.mark.parametrize("g", [2, 4])
def test_grouped_epilogue_lane_algebra(g):
    """Host model of the kGrouped epilogue of csrc/gemm_sm100.cu (recursive halving across the lanes of a run): for one
    64-column chunk of a tile (two 32-column slabs h = 0, 1) every 16-byte piece of every pooled staging row is written
    exactly once, by a lane of the right group, with the columns the TMA store expects there, and the value a lane keeps is
    the sum over exactly the g accumulator rows of its group."""
    rows = 128 // g
    written = {}
    for et in range(128):                                  # accumulator row = TMEM lane = epilogue thread
        lane = et % 32
        odd, hi2 = lane & 1, (lane >> 1) & 1
        for h in range(2):                                 # slab of 32 columns inside the 64-column chunk
            if g == 2:
                cols = [h * 32 + odd * 16 + k for k in range(16)]
                prow, pieces = et >> 1, [h * 4 + odd * 2, h * 4 + odd * 2 + 1]
                members = {et, et ^ 1}                      # shfl.xor 1
            else:
                cols = [h * 32 + odd * 16 + hi2 * 8 + k for k in range(8)]
                prow, pieces = et >> 2, [h * 4 + odd * 2 + hi2]
                members = {et, et ^ 1, et ^ 2, et ^ 3}      # shfl.xor 1, then shfl.xor 2
            assert {m // g for m in members} == {prow}, "a lane may only add rows of its own run"
            assert len(members) == g
            for i, piece in enumerate(pieces):
                key = (prow, piece)
                assert key not in written, "staging piece written twice"
                written[key] = cols[8 * i:8 * i + 8]
    assert len(written) == rows * 8                        # 64 columns = 8 pieces of 8 bf16 per pooled row
    for (prow, piece), cols in written.items():
        assert cols == list(range(piece * 8, piece * 8 + 8)), "piece p of a staging row holds columns [8p, 8p + 8)"
        assert 0 <= prow < rows
