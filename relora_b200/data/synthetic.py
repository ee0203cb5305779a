"""Synthetic token source: deterministic random ids of the benchmark shape.

Used by ``bench.py`` (BASELINE.json: "synthetic token ids / random-init weights"), by the CPU/gloo
plumbing config and by tests.  Sequence ``i`` is a pure function of ``(seed, i)`` so any rank can
produce any sample with no state (resume = index offset).
"""
from __future__ import annotations

import json
import os
from typing import Optional

import torch
from torch.utils.data import Dataset

__all__ = ["SyntheticTokens", "write_synthetic_hf_dataset"]


class SyntheticTokens(Dataset):
    """Ids are uniform over ``[0, vocab_size - 1)``: the last row is the embedding's padding row under the reference configs
    (``pad_token_id = -1``), which real tokenised text never contains -- and a sequence *starting* with it keeps an all-zero
    residual row whose RMSNorm backward (gain 1/sqrt(eps) per norm) overflows in deep models (bench.py:make_tokens)."""

    def __init__(self, n_sequences: int, seq_len: int, vocab_size: int, seed: int = 0):
        self.n, self.seq_len, self.vocab_size, self.seed = n_sequences, seq_len, vocab_size, seed

    def __len__(self):
        return self.n

    def __getitem__(self, i: int):
        g = torch.Generator()
        g.manual_seed((self.seed * 0x9E3779B1 + int(i) * 0x85EBCA77 + 12345) & 0x7FFFFFFFFFFFFFFF)
        return {"input_ids": torch.randint(0, max(1, self.vocab_size - 1), (self.seq_len,), generator=g, dtype=torch.long)}

    def shard(self, rank: int, world_size: int) -> "SyntheticTokens":
        """Contiguous shard, like ``split_dataset_by_node``."""
        per = self.n // world_size
        out = _Offset(self, rank * per, per)
        return out


class _Offset(Dataset):
    def __init__(self, base, start, n):
        self.base, self.start, self.n = base, start, n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return self.base[self.start + i]


def write_synthetic_hf_dataset(path: str, n_train: int, n_val: int, seq_len: int, vocab_size: int,
                               seed: int = 0, tokenizer: str = "t5-base") -> str:
    """Materialise a HF ``DatasetDict`` directory (+ ``args.json``) of random ids — what the
    reference trainer needs on disk (``torchrun_main.py:431-462``)."""
    import datasets

    g = torch.Generator().manual_seed(seed)
    train = torch.randint(0, max(1, vocab_size - 1), (n_train, seq_len), generator=g).tolist()
    val = torch.randint(0, max(1, vocab_size - 1), (n_val, seq_len), generator=g).tolist()
    dd = datasets.DatasetDict({
        "train": datasets.Dataset.from_dict({"input_ids": train}),
        "validation": datasets.Dataset.from_dict({"input_ids": val}),
    })
    dd.save_to_disk(path)
    with open(os.path.join(path, "args.json"), "w") as f:
        json.dump({"tokenizer": tokenizer, "sequence_length": seq_len, "dataset": "synthetic", "vocab_size": vocab_size}, f, indent=4)
    return path
