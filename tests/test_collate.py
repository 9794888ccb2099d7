"""Collator contract: ps_slm_b200.collate against the UNMODIFIED reference collator / dynamic
batcher where the reference tree is present, and against the contract's properties everywhere."""
import types

import pytest
import torch

from oracle import ref_loader as R

PAD, EOS, IGN, SP = 151643, 151643, -100, 151665


def _samples(seed, n, train):
    import ps_slm_b200.collate as C
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(n):
        p = torch.randint(0, 1000, (int(torch.randint(5, 30, (1,), generator=g)),), generator=g).tolist()
        p[len(p) // 2] = SP
        T = int(torch.randint(3, 40, (1,), generator=g))
        feats = torch.randn(T, 8, generator=g)
        tgt = torch.randint(0, 1000, (int(torch.randint(1, 12, (1,), generator=g)),), generator=g).tolist() if train else None
        out.append(C.build_sample(p, feats, key=f"utt{i}", target=f"t{i}", gt=f"g{i}", target_ids=tgt, eos_token_id=EOS))
    return out


@pytest.mark.parametrize("train", [True, False])
def test_collate_contract(train):
    import ps_slm_b200.collate as C
    s = _samples(1, 6, train)
    b = C.collate(s, PAD, IGN, inference_mode=not train, pin_memory=False)
    S = max(x["input_ids"].numel() for x in s)
    assert b["input_ids"].shape == (6, S) and b["input_ids"].dtype == torch.long and b["attention_mask"].dtype == torch.bool
    assert b["input_feature_length"].dtype == torch.long and b["GT"] == [f"g{i}" for i in range(6)]
    assert (b["input_ids"] == SP).sum(1).tolist() == [1] * 6                    # exactly one <speech> per row
    for i, x in enumerate(s):
        n = x["input_ids"].numel()
        if train:                                                              # right padding, labels -100 on prompt and pad
            assert torch.equal(b["input_ids"][i, :n], x["input_ids"]) and (b["input_ids"][i, n:] == PAD).all()
            assert b["attention_mask"][i, :n].all() and not b["attention_mask"][i, n:].any()
            assert torch.equal(b["labels"][i, :n], x["labels"]) and (b["labels"][i, n:] == IGN).all()
            assert int(b["labels"][i, n - 1]) == EOS
        else:                                                                  # left padding
            assert torch.equal(b["input_ids"][i, S - n:], x["input_ids"]) and (b["input_ids"][i, :S - n] == PAD).all()
            assert b["attention_mask"][i, S - n:].all() and not b["attention_mask"][i, :S - n].any()
    assert ("labels" in b) == train and ("keys" in b) == (not train)
    T = int(b["input_feature_length"].max())
    assert b["input_features"].shape == (6, T, 8)
    assert all(float(b["input_features"][i, int(b["input_feature_length"][i]):].abs().sum()) == 0 for i in range(6))


@pytest.mark.skipif(not R.available(), reason="reference tree not present")
@pytest.mark.parametrize("train", [True, False])
@pytest.mark.parametrize("seed", range(3))
def test_collate_matches_reference(train, seed):
    import ps_slm_b200.collate as C
    D = R.load_dataset_module()
    fake = types.SimpleNamespace(inference_mode=not train,
                                 tokenizer=types.SimpleNamespace(pad_token_id=PAD, default_ignore_token=IGN),
                                 dataset_config=types.SimpleNamespace(encoder="sensevoice"))
    fake.pad = types.MethodType(D.MultiTaskDataset.pad, fake)
    s = _samples(10 + seed, 5, train)
    ref = D.MultiTaskDataset.collator(fake, s)
    got = C.collate(s, PAD, IGN, inference_mode=not train, pin_memory=False)
    assert set(ref) == set(got)
    for k in ref:
        if isinstance(ref[k], torch.Tensor):
            assert ref[k].dtype == got[k].dtype and torch.equal(ref[k], got[k]), k
        else:
            assert ref[k] == got[k], k


@pytest.mark.skipif(not R.available(), reason="reference tree not present")
@pytest.mark.parametrize("budget", [60, 200, 1000])
def test_dynamic_batcher_matches_reference(budget):
    import functools

    import ps_slm_b200.collate as C
    D = R.load_dataset_module()
    s = _samples(5, 40, True)

    class DS(torch.utils.data.IterableDataset):
        collator = None

        def __iter__(self):
            return iter(s)

        def __len__(self):
            return len(s)
    ref = list(D.MultiTaskDynamicBatchDataset(DS(), functools.partial(D.window_class, max_frame_length=budget, ds_rate=2)))
    got = list(C.frame_budget_batches(s, budget, ds_rate=2))
    assert [[x["key"] for x in w] for w in ref] == [[x["key"] for x in w] for w in got]
    assert sum(len(w) for w in got) == len(s)
