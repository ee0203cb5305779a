"""Weighted mixture of datasets.

Behavioural target: ``megatron_dataset/blendable_dataset.py:27-79`` of the reference.  The mixture has as many samples as
its parts together; sample ``i`` of the mixture is sample ``sample_of[i]`` of part ``part_of[i]``.  The assignment is the
greedy "largest deficit first" schedule of the native ``build_blending_indices`` (csrc/data_helpers.cpp), which keeps the
running share of every part as close as possible to its normalised weight.
"""
from __future__ import annotations

import time
from typing import Sequence

import numpy as np
from torch.utils.data import Dataset

from ...obs import logger
from . import helpers_build

__all__ = ["BlendableDataset"]

_MAX_PARTS = 254  # the part id is stored in one byte


class BlendableDataset(Dataset):
    def __init__(self, datasets: Sequence[Dataset], weights: Sequence[float]):
        if len(datasets) != len(weights):
            raise AssertionError("one weight per dataset")
        if len(datasets) > _MAX_PARTS:
            raise AssertionError(f"at most {_MAX_PARTS} datasets can be blended")
        share = np.asarray(weights, dtype=np.float64)
        total = share.sum()
        if not total > 0.0:
            raise AssertionError("the blend weights must sum to a positive value")
        share = share / total

        self.datasets = list(datasets)
        self.size = int(sum(len(part) for part in self.datasets))
        # reference attribute names kept: checkpoints / notebooks poke at them
        self.dataset_index = np.zeros(self.size, dtype=np.uint8)
        self.dataset_sample_index = np.zeros(self.size, dtype=np.int64)
        began = time.time()
        helpers_build.load().build_blending_indices(self.dataset_index, self.dataset_sample_index, share, len(self.datasets), self.size,
                                                    False)
        took = time.time() - began
        if took > 5.0:
            logger.info(f"blend schedule for {self.size} samples built in {took:.1f} s")

    def __len__(self) -> int:
        return self.size

    def __getitem__(self, i: int):
        if i >= self.size or i < -self.size:
            # upstream tolerates an overrun of the sampler by wrapping around; keep that, but say so
            wrapped = i % self.size
            logger.warning(f"blend index {i} is outside [0, {self.size}); using {wrapped}")
            i = wrapped
        part = self.datasets[self.dataset_index[i]]
        return part[self.dataset_sample_index[i]]
