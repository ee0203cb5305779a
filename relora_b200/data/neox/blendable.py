"""Weighted interleaving of several datasets (reference ``megatron_dataset/blendable_dataset.py:27-79``); the
per-sample (dataset, index) assignment comes from the native ``build_blending_indices``."""
from __future__ import annotations

import time

import numpy as np
import torch

from . import helpers_build

__all__ = ["BlendableDataset"]


class BlendableDataset(torch.utils.data.Dataset):
    def __init__(self, datasets, weights):
        self.datasets = datasets
        n = len(datasets)
        assert n == len(weights) and n < 255
        self.size = sum(len(d) for d in datasets)
        w = np.array(weights, dtype=np.float64)
        assert w.sum() > 0.0
        w /= w.sum()
        t0 = time.time()
        self.dataset_index = np.zeros(self.size, dtype=np.uint8)
        self.dataset_sample_index = np.zeros(self.size, dtype=np.int64)
        helpers_build.load().build_blending_indices(self.dataset_index, self.dataset_sample_index, w, n, self.size, False)
        if time.time() - t0 > 5.0:
            print(f"> elapsed time for building blendable dataset indices: {time.time() - t0:.2f} (sec)")

    def __len__(self):
        return self.size

    def __getitem__(self, idx):
        try:
            return self.datasets[self.dataset_index[idx]][self.dataset_sample_index[idx]]
        except IndexError:
            new_idx = idx % len(self)
            print(f"WARNING: Got index out of bounds error with index {idx} - taking modulo of index instead ({new_idx})")
            return self[new_idx]
