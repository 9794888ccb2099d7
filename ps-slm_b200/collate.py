"""Collator contract feeding the bridge, host side.

Mirrors ``MultiTaskDataset.collator`` (Multitask/dataset/speech_dataset_large.py:240-305), the sample layout
built at :151-186 and the frame-budget dynamic batcher (``MultiTaskDynamicBatchDataset`` :307-330 with
``window_class`` :333-338), so that batches handed to ``slam_model_asr.forward/generate`` have exactly the
reference's keys, dtypes and padding:

* training  : ``input_ids = prompt ⊕ target ⊕ eos``, ``labels = -100`` on the prompt, RIGHT padding with
  ``pad_token_id`` / False / -100;
* inference : prompt only, LEFT padding; ``keys`` / ``targets`` lists ride along;
* ``input_features`` are zero-padded to the longest utterance, ``input_feature_length`` is int64.

The batch tensors are built directly in pinned host memory (one allocation per key) so the training loop's
``.to(device, non_blocking=True)`` is a true asynchronous copy.
"""
from typing import Dict, Iterable, Iterator, List, Optional, Sequence

import torch


def build_sample(prompt_ids: Sequence[int], input_features: torch.Tensor, key: str = "", target: str = "", gt: str = "",
                 target_ids: Optional[Sequence[int]] = None, eos_token_id: Optional[int] = None,
                 ignore_id: int = -100) -> Dict:
    """One sample as ``MultiTaskDataset.__iter__`` yields it (:162-186).  ``target_ids`` given → training sample."""
    prompt = torch.as_tensor(list(prompt_ids), dtype=torch.long)
    sample = {"input_features": input_features, "input_feature_length": int(input_features.shape[0]), "key": key,
              "target": target, "GT": gt}
    if target_ids is not None:
        tgt = torch.as_tensor(list(target_ids) + [eos_token_id], dtype=torch.long)
        ids = torch.cat([prompt, tgt])
        labels = ids.clone()
        labels[:prompt.numel()] = ignore_id
        sample["labels"] = labels
    else:
        ids = prompt
    sample["input_ids"] = ids
    sample["attention_mask"] = ids.ge(-1)
    return sample


def _pinned(shape, dtype, fill, pin):
    t = torch.empty(shape, dtype=dtype, pin_memory=pin)
    t.fill_(fill)
    return t


def collate(samples: List[Dict], pad_token_id: int, ignore_id: int = -100, inference_mode: bool = False,
            pin_memory: Optional[bool] = None) -> Dict:
    """Drop-in for ``MultiTaskDataset.collator`` on the SenseVoice branch (:240-305)."""
    assert samples is not None
    pin = torch.cuda.is_available() if pin_memory is None else pin_memory
    B = len(samples)
    S = max(s["input_ids"].shape[0] for s in samples)
    input_ids = _pinned((B, S), torch.long, pad_token_id, pin)
    attention_mask = _pinned((B, S), torch.bool, False, pin)
    labels = None if inference_mode else _pinned((B, S), torch.long, ignore_id, pin)
    for b, s in enumerate(samples):
        n = s["input_ids"].shape[0]
        sl = slice(S - n, S) if inference_mode else slice(0, n)          # left padding at inference, right in training
        input_ids[b, sl] = s["input_ids"]
        attention_mask[b, sl] = s["attention_mask"]
        if labels is not None:
            labels[b, sl] = s["labels"]
    t_max = max(s["input_features"].size(0) for s in samples)
    feat_dim = samples[0]["input_features"].shape[1:]
    input_features = _pinned((B, t_max) + tuple(feat_dim), samples[0]["input_features"].dtype, 0.0, pin)
    for b, s in enumerate(samples):
        input_features[b, :s["input_features"].size(0)] = s["input_features"]
    result = {
        "input_ids": input_ids,
        "attention_mask": attention_mask,
        "input_features": input_features,
        "input_feature_length": torch.tensor([s["input_feature_length"] for s in samples], dtype=torch.long),
        "GT": [s["GT"] for s in samples],
    }
    if inference_mode:
        result["keys"] = [s["key"] for s in samples]
        result["targets"] = [s["target"] for s in samples]
    else:
        result["labels"] = labels
    return result


def spliced_length(sample: Dict, ds_rate: int) -> int:
    """Upper bound of the sample's length after the splice, as ``window_class`` counts it (:336)."""
    return len(sample["input_ids"]) + (sample["input_feature_length"] // ds_rate) - 1


def frame_budget_batches(samples: Iterable[Dict], max_frame_length: int, ds_rate: int = 1) -> Iterator[List[Dict]]:
    """Dynamic batcher of ``MultiTaskDynamicBatchDataset`` (:307-330): a sample joins the current window unless
    ``(len(window) + 1) · max(spliced length)`` would exceed ``max_frame_length`` (``window_class`` :333-338); note
    the reference's predicate returns True for an EMPTY buffer, so the very first sample opens a new window after
    the (empty) current one — reproduced by never yielding empty windows."""
    buf: List[Dict] = []
    for elem in samples:
        if len(buf) == 0:
            flush = True                                                # window_class(elem, []) is True (:335)
        else:
            mx = max(spliced_length(elem, ds_rate), max(spliced_length(s, ds_rate) for s in buf))
            flush = (len(buf) + 1) * mx > max_frame_length
        if not flush:
            buf.append(elem)
        else:
            if len(buf) > 0:
                yield buf
            buf = [elem]
    if len(buf) > 0:
        yield buf
