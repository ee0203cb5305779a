"""Pre-tokenised HF-dataset pipeline: tokenise + chunk, rank sharding, O(1)-resume loader.

Parity target: reference ``peft_pretraining/dataloader.py`` (``tokenize_and_chunk`` ``:57-124``,
``SkipBatchSampler`` ``:128-147``, ``SkipDataLoader`` ``:150-170``, ``PreprocessedIterableDataset``
``:13-54``) and the dataset handling in ``torchrun_main.py:431-462, 718-741``.

Upstream resumes by iterating the loader and throwing batches away (tokenises/loads every skipped
batch).  Here a batch is a pure function of its index, so resuming is an offset into the index
stream; batches are produced into pinned host memory so the trainer's H2D copy is asynchronous.
"""
from __future__ import annotations

import itertools
import json
import os
from itertools import chain
from typing import Dict, Iterator, Optional

import torch
from torch.utils.data import BatchSampler, DataLoader, IterableDataset, SequentialSampler, get_worker_info

from ..obs import logger

__all__ = [
    "tokenize_and_chunk",
    "SkipBatchSampler",
    "SkipDataLoader",
    "PreprocessedIterableDataset",
    "load_pretokenized",
    "shard_for_rank",
    "collate_input_ids",
    "check_dataset_size",
]


def tokenize_and_chunk(tokenizer, dataset, text_field: str, sequence_length: int, num_cpu: Optional[int] = None):
    """Tokenise ``text + eos``, concatenate everything and cut into ``sequence_length`` blocks.

    The tail shorter than a block is dropped and ``attention_mask`` is not stored (there is never
    padding).  Works on a ``datasets.Dataset`` / ``DatasetDict`` (parallel map) or an iterable one.
    """
    map_kw = {} if isinstance(dataset, IterableDataset) or num_cpu in (None, 0, 1) else {"num_proc": num_cpu}
    eos = tokenizer.eos_token or ""

    def _tok(batch):
        return tokenizer([t + eos for t in batch[text_field]])

    n_before = len(dataset) if hasattr(dataset, "__len__") else None
    tokenized = dataset.map(_tok, batched=True, remove_columns=[text_field], **map_kw)
    if n_before is not None:
        assert len(tokenized) == n_before
    logger.info("Tokenization finished")

    def _group(batch):
        flat = {k: list(chain.from_iterable(batch[k])) for k in batch.keys()}
        n = len(flat["input_ids"])
        if n >= sequence_length:
            n = (n // sequence_length) * sequence_length
        return {
            k: [v[i : i + sequence_length] for i in range(0, n, sequence_length)]
            for k, v in flat.items()
            if k != "attention_mask"
        }

    cols = tokenized.column_names
    if isinstance(cols, dict):  # DatasetDict
        has_mask = any("attention_mask" in c for c in cols.values())
    else:
        has_mask = cols is not None and "attention_mask" in cols
    chunked = tokenized.map(_group, batched=True, remove_columns=["attention_mask"] if has_mask else None, **map_kw)
    logger.info("Chunking finished")
    return chunked


class PreprocessedIterableDataset(IterableDataset):
    """On-the-fly tokenisation with padding to ``max_length`` (kept for API parity; the trainer uses
    the pre-tokenised path)."""

    def __init__(self, data, tokenizer, batch_size: int, max_length: int):
        super().__init__()
        self.data, self.tokenizer, self.batch_size, self.max_length = data, tokenizer, batch_size, max_length

    def __iter__(self):
        info = get_worker_info()
        it = iter(self.data) if info is None else itertools.islice(self.data, info.id, None, info.num_workers)
        batch = []
        for ex in it:
            batch.append(self.tokenizer(ex["text"], max_length=self.max_length, truncation=True, padding="max_length", return_tensors="pt"))
            if len(batch) == self.batch_size:
                yield self._format(batch)
                batch = []
        if batch:
            yield self._format(batch)

    @staticmethod
    def _format(batch):
        return {
            "input_ids": torch.stack([b["input_ids"].squeeze(0) for b in batch]),
            "attention_mask": torch.stack([b["attention_mask"].squeeze(0) for b in batch]),
        }


class SkipBatchSampler(BatchSampler):
    """Wraps a batch sampler and starts at batch ``skip_batches`` without materialising the skipped
    index lists when the inner sampler is sequential."""

    def __init__(self, batch_sampler, skip_batches: int = 0):
        self.batch_sampler = batch_sampler
        self.skip_batches = skip_batches

    def __iter__(self):
        inner = self.batch_sampler
        seq = isinstance(getattr(inner, "sampler", None), SequentialSampler)
        if seq and self.skip_batches > 0:
            n, bs = len(inner.sampler), inner.batch_size
            start = self.skip_batches * bs
            for lo in range(start, n, bs):
                idx = list(range(lo, min(lo + bs, n)))
                if len(idx) < bs and inner.drop_last:
                    return
                yield idx
            return
        for i, samples in enumerate(inner):
            if i >= self.skip_batches:
                yield samples

    @property
    def total_length(self):
        return len(self.batch_sampler)

    def __len__(self):
        return max(0, len(self.batch_sampler) - self.skip_batches)


def collate_input_ids(samples) -> Dict[str, torch.Tensor]:
    """``default_data_collator`` restricted to what the trainer consumes: int64 ``input_ids``."""
    first = samples[0]["input_ids"]
    if isinstance(first, torch.Tensor):
        ids = torch.stack([s["input_ids"] for s in samples])
    else:
        import numpy as np

        ids = torch.from_numpy(np.stack([np.asarray(s["input_ids"]) for s in samples]))
    return {"input_ids": ids.long()}


class SkipDataLoader(DataLoader):
    """``DataLoader`` whose first ``skip_batches`` batches are never produced (index-level skip)."""

    def __init__(self, dataset, skip_batches: int = 0, **kwargs):
        self.skip_batches = skip_batches
        if skip_batches and "batch_sampler" not in kwargs and not isinstance(dataset, IterableDataset) and not kwargs.get("shuffle", False):
            bs = kwargs.pop("batch_size", 1)
            drop = kwargs.pop("drop_last", False)
            kwargs.pop("shuffle", None)
            kwargs.pop("sampler", None)
            inner = BatchSampler(SequentialSampler(dataset), bs, drop)
            kwargs["batch_sampler"] = SkipBatchSampler(inner, skip_batches)
            self._index_skip = True
        else:
            self._index_skip = False
        super().__init__(dataset, **kwargs)

    def __iter__(self) -> Iterator:
        if self._index_skip or not self.skip_batches:
            yield from super().__iter__()
            return
        for i, batch in enumerate(super().__iter__()):
            if i >= self.skip_batches:
                yield batch


def load_pretokenized(path: str, seed: int = 0):
    """``datasets.load_from_disk`` + torch format; returns ``(train, validation, args_json)``.

    Shuffles the training split only when ``seed != 0`` (upstream's backward-compat condition,
    ``torchrun_main.py:438-440``).
    """
    import datasets

    dd = datasets.load_from_disk(path)
    # plain python rows: the torch / numpy formatters of `datasets` import torchvision (often broken or absent on
    # training boxes); the collate function turns the id lists into one int64 tensor per batch
    dd.set_format(type=None, columns=["input_ids"])
    train = dd["train"]
    if seed != 0:
        train = train.shuffle(seed=seed)
    val = dd["validation"] if "validation" in dd else None
    with open(os.path.join(path, "args.json")) as f:
        prep = json.load(f)
    return train, val, prep


def check_dataset_size(n_sequences: int, max_length: int, total_batch_size: int, num_training_steps: int, parity_quirks: bool = True):
    """Upstream compares *sequences needed* against *tokens available* (``torchrun_main.py:447-450``);
    with ``parity_quirks=False`` the comparison is sequences vs sequences."""
    need = total_batch_size * num_training_steps
    have = n_sequences * max_length if parity_quirks else n_sequences
    if have < need:
        unit = "tokens" if parity_quirks else "sequences"
        raise ValueError(f"Dataset only has {have} {unit}, but we need at least {need}")


def shard_for_rank(dataset, rank: int, world_size: int):
    """``datasets.distributed.split_dataset_by_node`` (contiguous shard per rank)."""
    if world_size == 1:
        return dataset
    import datasets.distributed

    return datasets.distributed.split_dataset_by_node(dataset, rank=rank, world_size=world_size)
