"""CPU: the oracle restatement against the golden vectors frozen from the real reference."""
import numpy as np
import pytest
import torch

from oracle import tasu_oracle as O
from conftest import expand_posterior

SP, PAD = 151665, 151643


def _t(a):
    return torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize("fn", ["loop", "vec"])
def test_psd_golden(golden, fn):
    g = golden["psd"]

    def run(feats, lens, post, blank=0, thr=0.9):
        if fn == "loop":
            return O.psd_loop(feats, lens, post, blank, thr)
        f, l, _ = O.psd_vec(feats, lens, post, blank, thr)
        return f, l

    p1 = _t(g["t1_post"])
    f, l = run(p1, _t(g["t1_lens"]), p1)
    assert l.tolist() == g["t1_ref_lens"].tolist() == [6]
    np.testing.assert_allclose(f.numpy(), g["t1_ref_feats"], rtol=1e-6, atol=1e-7)
    # the SURVEY decision table: kept candidates have argmax [0,5,7,0,5,0], scores [.6,.2,.05,.89,0,.8999]
    assert f[0].argmax(-1).tolist() == [0, 5, 7, 0, 5, 0]
    np.testing.assert_allclose(f[0, :, 0].numpy(), [.6, .2, .05, .89, 0, .8999], atol=1e-6)

    p2, lens2 = _t(g["t2_post"]), _t(g["t2_lens"])
    f, l = run(p2, lens2, p2)
    assert l.tolist() == g["t2_ref_lens"].tolist()
    np.testing.assert_allclose(f.numpy(), g["t2_ref_feats"], rtol=1e-6, atol=1e-7)
    f, l = run(_t(g["t3_feats"]), lens2, p2.log())
    assert l.tolist() == g["t3_ref_lens"].tolist()
    np.testing.assert_allclose(f.numpy(), g["t3_ref_feats"], rtol=1e-5, atol=1e-6)
    f, l = run(p2, lens2, p2, 3, 0.5)
    assert l.tolist() == g["t4_ref_lens"].tolist()
    np.testing.assert_allclose(f.numpy(), g["t4_ref_feats"], rtol=1e-6, atol=1e-7)
    f, l = run(p2, torch.zeros(5, dtype=torch.long), p2)
    assert list(f.shape) == g["t5_ref_shape"].tolist() and l.tolist() == g["t5_ref_lens"].tolist()


def test_psd_golden_full_vocab(golden):
    g = golden["psd"]
    V = 25055
    p = expand_posterior(g["t6_lab"], g["t6_alt"], g["t6_w1"], g["t6_w2"], V)
    f, l, _ = O.psd_vec(p, _t(g["t6_lens"]), p)
    assert l.tolist() == g["t6_ref_lens"].tolist()
    np.testing.assert_allclose(f.numpy()[:, :, g["t6_cols"]], g["t6_ref_feats_cols"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(f.double().sum(-1).numpy(), g["t6_ref_rowsum"], rtol=1e-6)


def _ids_list(g):
    flat, lens = g["ids_flat"].tolist(), g["ids_len"].tolist()
    out, o = [], 0
    for n in lens:
        out.append(flat[o:o + n])
        o += n
    return out


def test_sim_golden(golden):
    g = golden["sim"]
    ids = _ids_list(g)
    V = 25055
    p, l = O.sim_posterior_clean(ids, V)
    assert l.tolist() == g["clean_ref_lens"].tolist()
    assert np.array_equal(p.argmax(-1).numpy(), g["clean_ref_argmax"])
    assert np.array_equal(p.sum(-1).numpy(), g["clean_ref_sum"])
    for name, ip in (("n0", 0.0), ("n1", 0.1)):
        torch.manual_seed(1234)
        p, l = O.sim_posterior_noise(ids, V, 0, insert_prob=ip)
        assert l.tolist() == g[f"{name}_ref_lens"].tolist()
        am = p.argmax(-1)
        assert np.array_equal(am.numpy(), g[f"{name}_ref_argmax"])
        hot = p.gather(-1, am.unsqueeze(-1)).squeeze(-1)
        assert np.array_equal(hot.numpy(), g[f"{name}_ref_hot"])          # bit-exact fp32
        other = torch.where(am == 1, 2, 1)
        base = p.gather(-1, other.unsqueeze(-1)).squeeze(-1)
        assert np.array_equal(base.numpy(), g[f"{name}_ref_base"])


def test_projector_golden(golden):
    g = golden["projector"]
    y = O.projector_linear_silu(_t(g["linear_silu_x"]), _t(g["linear_silu_p_norm.weight"]), _t(g["linear_silu_p_norm.bias"]),
                                _t(g["linear_silu_p_ffn.0.weight"]), _t(g["linear_silu_p_ffn.0.bias"]),
                                _t(g["linear_silu_p_ffn.2.weight"]), _t(g["linear_silu_p_ffn.2.bias"]))
    np.testing.assert_allclose(y.numpy(), g["linear_silu_ref_y"], rtol=1e-5, atol=1e-6)
    y = O.projector_concat(_t(g["linear_x"]), 2, _t(g["linear_p_linear1.weight"]), _t(g["linear_p_linear1.bias"]),
                           _t(g["linear_p_linear2.weight"]), _t(g["linear_p_linear2.bias"]))
    np.testing.assert_allclose(y.numpy(), g["linear_ref_y"], rtol=1e-5, atol=1e-6)
    y = O.projector_linear(_t(g["simple_linear_x"]), 3, _t(g["simple_linear_p_map.weight"]), _t(g["simple_linear_p_map.bias"]))
    np.testing.assert_allclose(y.numpy(), g["simple_linear_ref_y"], rtol=1e-5, atol=1e-6)


MERGE_CASES = ["right", "left", "nopad", "zero", "single"]


@pytest.mark.parametrize("name", MERGE_CASES)
def test_merge_golden(golden, name):
    g = golden["merge"]
    lab = _t(g[f"{name}_lab"]) if f"{name}_lab" in g.files else None
    e, m, l, p, f = O.merge(_t(g[f"{name}_af"]), _t(g[f"{name}_M"]), _t(g[f"{name}_emb"]), _t(g[f"{name}_ids"]),
                            _t(g[f"{name}_att"]), lab, SP, PAD)
    assert np.array_equal(e.numpy(), g[f"{name}_ref_emb"])
    assert np.array_equal(m.numpy(), g[f"{name}_ref_mask"]) and m.dtype == torch.bool
    assert np.array_equal(p.numpy(), g[f"{name}_ref_pos"])
    assert np.array_equal(f.numpy(), g[f"{name}_ref_ids"])
    if lab is None:
        assert l is None
    else:
        assert np.array_equal(l.numpy(), g[f"{name}_ref_lab"])


@pytest.mark.parametrize("name", ["err_both", "err_rpad1"])
def test_merge_golden_errors(golden, name):
    g = golden["merge"]
    ids, att, M = _t(g[name + "_ids"]), _t(g[name + "_att"]).bool(), _t(g[name + "_M"])
    with pytest.raises(ValueError):
        O.merge(torch.zeros(ids.shape[0], int(M.max()), 8), M, torch.zeros(*ids.shape, 8), ids, att, None, SP, PAD)


@pytest.mark.parametrize("name", ["infer_cross_attn", "infer_voca_trans"])
def test_oracle_reproduces_model_golden(golden, name):
    """CPU: the oracle restatements of the cross-attention projector (projector.py:104-126) and of the voca_trans branch
    (ps-slm.py:485-516) reproduce what the reference handed to the LLM (tests/golden/model.npz; the voca_trans case comes
    from the reference source with its one-line UnboundLocalError fix, oracle/ref_loader.py:load_voca_fixed)."""
    import fakes as F
    import torch
    from oracle import tasu_oracle as O
    import ps_slm_b200.projector as P                       # only as a parameter container with the reference's names/init
    g = golden["model"]
    inp = F.build_inputs(name)
    encoder, llm, projector, tok, train_config, model_config = F.build_parts(inp, P.PROJECTORS[inp["proj"]])
    wsum = sum(float(p.detach().double().sum()) for m in (encoder, llm, projector) for p in m.parameters())
    assert abs(wsum - float(g[f"{name}_wsum"])) < 1e-6 * max(1.0, abs(wsum))
    batch = inp["batch"]
    with torch.no_grad():
        raw, raw_lens = encoder.encoder(torch.zeros(batch["input_features"].shape[0], batch["input_features"].shape[1] + 4, 1),
                                        batch["input_feature_length"] + 4)
        lens = torch.clamp(raw_lens.long() - 4, min=0)
        table = llm.emb.weight
        if name == "infer_cross_attn":
            post = torch.softmax(encoder.ctc.ctc_lo(raw), -1)[:, 4:]
            outs, feat_len, _ = O.psd_vec(post, lens, post, 0, 0.9)
            proj = O.projector_ctcca(outs, table, projector.W_q.weight, projector.n_heads)
        else:
            proj, feat_len = O.voca_trans(raw[:, 4:], lens, projector.map.weight, projector.map.bias, projector.k, table,
                                          True, False)
        emb, mask, _, _, _ = O.merge(proj, feat_len, table[batch["input_ids"]], batch["input_ids"], batch["attention_mask"],
                                     None, F.SPEECH_ID, F.PAD_ID)
    assert np.array_equal(mask.numpy(), g[f"{name}_ref_mask"])
    ref = torch.from_numpy(g[f"{name}_ref_embeds"])
    assert emb.shape == ref.shape and torch.allclose(emb, ref, rtol=1e-4, atol=1e-5)
