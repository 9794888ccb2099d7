"""Freeze golden vectors by running the UNMODIFIED reference (container only).

    python oracle/make_golden.py        # writes tests/golden/*.npz

Test infrastructure only.  Every array named ``ref_*`` is an output of the real
reference functions loaded by oracle/ref_loader.py; everything else is the
input that produced it.  The files travel to the GPU box (where /root/reference
does not exist) and pin both the oracle restatement and the CUDA path.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader as R  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SP, PAD = 151665, 151643


def compact_posterior(B, T, V, seed, blank=0):
    """Deterministic peaky posterior described by a few integers per frame, so a
    V=25055 case fits in a tiny fixture: p = base/V + w1*e_i + w2*e_j."""
    g = np.random.RandomState(seed)
    lab = g.randint(1, V, size=(B, T))
    lab[g.rand(B, T) < 0.45] = blank
    for t in range(1, T):
        rep = g.rand(B) < 0.35
        lab[rep, t] = lab[rep, t - 1]
    alt = g.randint(0, V, size=(B, T))
    w1 = g.choice([0.55, 0.7, 0.86, 0.93, 0.97, 0.995], size=(B, T)).astype(np.float32)
    w2 = ((1.0 - w1) * g.choice([0.0, 0.5, 0.9], size=(B, T))).astype(np.float32)
    return lab.astype(np.int64), alt.astype(np.int64), w1, w2


def expand_posterior(lab, alt, w1, w2, V):
    B, T = lab.shape
    base = (1.0 - w1 - w2) / V
    p = np.repeat(base[..., None], V, axis=2).astype(np.float32)
    bi, ti = np.meshgrid(np.arange(B), np.arange(T), indexing="ij")
    p[bi, ti, lab] += w1
    p[bi, ti, alt] += w2
    return torch.from_numpy(p)


def golden_psd():
    out = {}
    # 1. the decision table of SURVEY.md §8a (threshold is strict '<' in fp32)
    ids = [0, 0, 5, 5, 5, 0, 7, 7, 0, 5, 0, 0]
    pb = [.95, .6, .1, .2, .3, .91, .05, .05, .89, 0, .9, .8999]
    V = 9
    p = torch.zeros(1, len(ids), V)
    for t, (i, q) in enumerate(zip(ids, pb)):
        if i == 0:
            p[0, t, 0] = q
            p[0, t, 1:] = (1 - q) / (V - 1) * 0.5   # keep blank the argmax
            p[0, t, 0] = q
        else:
            p[0, t, 0] = q
            p[0, t, i] = max(1 - q, q + 0.01)
    lens = torch.tensor([len(ids)])
    f, nl = R.ref_psd(p, lens, p, 0, 0.9)
    out.update(t1_post=p.numpy(), t1_lens=lens.numpy(), t1_ref_feats=f.numpy(), t1_ref_lens=nl.numpy())
    # 2. random small-V batch with ragged lengths incl. L=0, posterior == features
    torch.manual_seed(11)
    B, T, V = 5, 40, 13
    lab, alt, w1, w2 = compact_posterior(B, T, V, 5)
    p = expand_posterior(lab, alt, w1, w2, V)
    lens = torch.tensor([40, 0, 17, 1, 33])
    f, nl = R.ref_psd(p, lens, p, 0, 0.9)
    out.update(t2_post=p.numpy(), t2_lens=lens.numpy(), t2_ref_feats=f.numpy(), t2_ref_lens=nl.numpy())
    # 3. log-prob input, raw 16-d features (ctc_posterior=false path, ps-slm.py:518)
    feats = torch.randn(B, T, 16)
    lp = p.log()
    f, nl = R.ref_psd(feats, lens, lp, 0, 0.9)
    out.update(t3_feats=feats.numpy(), t3_ref_feats=f.numpy(), t3_ref_lens=nl.numpy())
    # 4. non-zero blank id + different threshold
    f, nl = R.ref_psd(p, lens, p, 3, 0.5)
    out.update(t4_ref_feats=f.numpy(), t4_ref_lens=nl.numpy())
    # 5. all-empty → [B,0,D]
    f, nl = R.ref_psd(p, torch.zeros(B, dtype=torch.long), p, 0, 0.9)
    out.update(t5_ref_shape=np.array(f.shape), t5_ref_lens=nl.numpy())
    # 6. full-width V=25055 (compact description; only a column sample of the pooled rows is stored)
    B, T, V = 3, 30, 25055
    lab, alt, w1, w2 = compact_posterior(B, T, V, 77)
    p = expand_posterior(lab, alt, w1, w2, V)
    lens = torch.tensor([30, 21, 9])
    f, nl = R.ref_psd(p, lens, p, 0, 0.9)
    cols = np.unique(np.concatenate([np.arange(0, V, 997), lab.reshape(-1), alt.reshape(-1), [0, 1, 2, 3, V - 4, V - 3, V - 2, V - 1]]))
    out.update(t6_lab=lab, t6_alt=alt, t6_w1=w1, t6_w2=w2, t6_lens=lens.numpy(), t6_cols=cols,
               t6_ref_feats_cols=f.numpy()[:, :, cols], t6_ref_lens=nl.numpy(),
               t6_ref_rowsum=f.double().sum(-1).numpy(), t6_ref_rowsq=(f.double() ** 2).sum(-1).numpy())
    np.savez_compressed(os.path.join(OUT, "psd.npz"), **out)


def golden_sim():
    out = {}
    V = 25055
    g = np.random.RandomState(1234)
    ids_list = []
    for _ in range(8):
        L = int(g.randint(35, 106))
        ids = g.randint(1, V, size=L)
        for t in range(1, L):
            if g.rand() < 0.05:
                ids[t] = ids[t - 1]
        ids_list.append([int(v) for v in ids])
    texts = [" ".join(map(str, i)) for i in ids_list]
    flat = np.concatenate([np.array(i) for i in ids_list])
    out.update(ids_flat=flat, ids_len=np.array([len(i) for i in ids_list]))
    p, l = R.ref_sim_clean(texts, V)
    out.update(clean_ref_lens=l.numpy(), clean_ref_argmax=p.argmax(-1).numpy(), clean_ref_sum=p.sum(-1).numpy())
    for name, ip in (("n0", 0.0), ("n1", 0.1)):
        torch.manual_seed(1234)
        p, l = R.ref_sim_noise(texts, V, 0, insert_prob=ip)
        # row descriptors recovered from the dense reference output
        am = p.argmax(-1)
        hot = p.gather(-1, am.unsqueeze(-1)).squeeze(-1)
        other = torch.where(am == 1, 2, 1)
        base = p.gather(-1, other.unsqueeze(-1)).squeeze(-1)
        out.update({f"{name}_ref_lens": l.numpy(), f"{name}_ref_argmax": am.numpy(),
                    f"{name}_ref_hot": hot.numpy(), f"{name}_ref_base": base.numpy(),
                    f"{name}_ref_rowsum": p.double().sum(-1).numpy()})
    np.savez_compressed(os.path.join(OUT, "sim.npz"), **out)


def golden_projector():
    out = {}
    torch.manual_seed(0)
    for kind, k, D in (("linear-silu", 1, 67), ("linear", 2, 24), ("simple_linear", 3, 20)):
        m = R.ref_projector(kind, D, 32, k)
        if kind == "linear-silu":      # non-trivial affine so the LayerNorm fold is exercised
            with torch.no_grad():
                m.norm.weight.uniform_(0.5, 1.5)
                m.norm.bias.uniform_(-0.2, 0.2)
                m.ffn[2].bias.uniform_(-0.1, 0.1)
        x = torch.softmax(torch.randn(3, 11, D) * 4, -1) if kind == "linear-silu" else torch.randn(3, 11, D)
        x[2, 7:] = 0                                   # zero-padded rows as produced by psd
        y = m(x)
        key = kind.replace("-", "_")
        out[f"{key}_x"] = x.numpy()
        out[f"{key}_ref_y"] = y.detach().numpy()
        for n, t in m.state_dict().items():
            out[f"{key}_p_{n}"] = t.numpy()
    np.savez_compressed(os.path.join(OUT, "projector.npz"), **out)


def golden_merge():
    out = {}
    cases = {
        "right": ([[1, 2, SP, 3, 4, 5], [6, SP, 7, PAD, PAD, PAD]], [[1] * 6, [1, 1, 1, 0, 0, 0]], [2, 4], True),
        "left": ([[1, 2, 3, SP, 4], [PAD, PAD, 6, SP, 7]], [[1] * 5, [0, 0, 1, 1, 1]], [3, 1], False),
        "nopad": ([[1, SP, 2], [3, SP, 4]], [[1] * 3, [1] * 3], [1, 3], True),
        "zero": ([[1, SP, 2], [3, SP, 4]], [[1] * 3, [1] * 3], [0, 2], True),
        "single": ([[1, SP, 2]], [[1] * 3], [5], False),
    }
    torch.manual_seed(5)
    H = 8
    for name, (ids, att, M, with_labels) in cases.items():
        ids = torch.tensor(ids)
        att = torch.tensor(att).bool()
        M = torch.tensor(M)
        B, S = ids.shape
        emb = torch.randn(B, S, H)
        af = torch.randn(B, max(int(M.max()), 1), H)
        lab = torch.randint(0, 1000, (B, S)) if with_labels else None
        e, m, l, p, f = R.ref_merge(af, M, emb, ids, att, lab, SP, PAD)
        out.update({f"{name}_ids": ids.numpy(), f"{name}_att": att.numpy(), f"{name}_M": M.numpy(),
                    f"{name}_emb": emb.numpy(), f"{name}_af": af.numpy(),
                    f"{name}_ref_emb": e.numpy(), f"{name}_ref_mask": m.numpy(),
                    f"{name}_ref_pos": p.numpy(), f"{name}_ref_ids": f.numpy()})
        if lab is not None:
            out[f"{name}_lab"] = lab.numpy()
            out[f"{name}_ref_lab"] = l.numpy()
    # error cases (reference raises ValueError): both-side padding; slot-count mismatch
    out["err_both_ids"] = np.array([[PAD, 1, SP], [3, SP, PAD]])
    out["err_both_att"] = np.array([[0, 1, 1], [1, 1, 0]])
    out["err_both_M"] = np.array([1, 1])
    # B == 1 is always treated as left padding (:774-775), so a right-padded single row miscounts slots
    out["err_rpad1_ids"] = np.array([[1, SP, 2, PAD]])
    out["err_rpad1_att"] = np.array([[1, 1, 1, 0]])
    out["err_rpad1_M"] = np.array([3])
    for name in ("err_both", "err_rpad1"):
        ids = torch.from_numpy(out[name + "_ids"]); att = torch.from_numpy(out[name + "_att"]).bool()
        M = torch.from_numpy(out[name + "_M"])
        try:
            R.ref_merge(torch.zeros(ids.shape[0], int(M.max()), H), M, torch.zeros(*ids.shape, H), ids, att, None, SP, PAD)
            raise SystemExit("reference did not raise for " + name)
        except ValueError as e:
            out[name + "_ref_error"] = np.array(str(e)[:40])
    np.savez_compressed(os.path.join(OUT, "merge.npz"), **out)


def golden_model():
    """Drive the UNMODIFIED reference ``slam_model_asr.forward/generate`` (ps-slm.py:411-677) with the
    fake encoder/LLM of tests/fakes.py and freeze what it hands to the LLM."""
    import contextlib
    import io
    sys.path.insert(0, os.path.join(os.path.dirname(OUT)))
    import fakes as F
    mod, pmod = R.load()
    out = {}
    for name in F.CASES:
        inp = F.build_inputs(name)
        cls = {"linear-silu": pmod.EncoderProjectorLinearSiLU, "linear": pmod.EncoderProjectorConcat,
               "cross-attention": pmod.EncoderProjectorCTCCA, "simple_linear": pmod.EncoderProjectorLinear}[inp["proj"]]
        encoder, llm, projector, tok, train_config, model_config = F.build_parts(inp, cls)
        # the shipped voca_trans branch raises UnboundLocalError: that one case runs the source with the one-line fix
        mod_c = R.load_voca_fixed() if inp["flags"].get("voca_trans") else mod
        m = mod_c.slam_model_asr.__new__(mod_c.slam_model_asr)
        torch.nn.Module.__init__(m)
        m.encoder, m.llm, m.encoder_projector, m.tokenizer = encoder, llm, projector, tok
        m.metric = "acc"
        m.train_config, m.model_config = train_config, model_config
        for k, v in inp["flags"].items():
            setattr(m, k, v)
        m.cross_attn = inp["proj"] == "cross-attention"       # ps-slm.py:229
        m.encoder_tokenizer = F.FakeCTCTokenizer()
        torch.manual_seed(4321)                         # RNG stream of the noisy simulator
        with contextlib.redirect_stdout(io.StringIO()):
            if inp["entry"] == "forward":
                res, acc = m(**inp["batch"])
                out[f"{name}_ref_loss"] = res.loss.detach().numpy()
                out[f"{name}_ref_acc"] = np.asarray(float(acc))
            else:
                m.generate(**inp["batch"])
        seen = llm.seen
        out[f"{name}_ref_embeds"] = seen["inputs_embeds"].detach().numpy()
        out[f"{name}_ref_mask"] = seen["attention_mask"].numpy()
        if seen.get("labels") is not None:
            out[f"{name}_ref_labels"] = seen["labels"].numpy()
        if seen.get("position_ids") is not None:
            out[f"{name}_ref_pos"] = seen["position_ids"].numpy()
        out[f"{name}_wsum"] = np.asarray(sum(float(p.double().sum()) for p in m.parameters()))
    np.savez_compressed(os.path.join(OUT, "model.npz"), **out)


def main():
    if not R.available():
        raise SystemExit("reference tree not present; golden vectors can only be regenerated in the build container")
    os.makedirs(OUT, exist_ok=True)
    golden_psd()
    golden_sim()
    golden_projector()
    golden_merge()
    golden_model()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
