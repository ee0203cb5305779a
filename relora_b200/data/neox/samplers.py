"""Batch-level distributed sampling with O(1) resume (reference ``megatron_dataset/samplers.py``).

``DistributedBatchSampler`` walks the global batch stream and hands each rank its slice; ``start_iter`` skips whole
batches *by index arithmetic* when the wrapped sampler is sequential (the reference iterates and discards)."""
from __future__ import annotations

import torch
from torch.utils import data

__all__ = ["DistributedBatchSampler", "RandomSampler"]


class RandomSampler(data.sampler.Sampler):
    """Epoch-seeded random sampler (``set_epoch``), with or without replacement."""

    def __init__(self, data_source, replacement=False, num_samples=None):
        self.data_source = data_source
        self.replacement = replacement
        self._num_samples = num_samples
        self.epoch = -1
        if self._num_samples is not None and replacement is False:
            raise ValueError("With replacement=False, num_samples should not be specified, since a random permute will be performed.")
        if not isinstance(self.num_samples, int) or self.num_samples <= 0:
            raise ValueError(f"num_samples should be a positive integer value, but got num_samples={self.num_samples}")
        if not isinstance(self.replacement, bool):
            raise ValueError(f"replacement should be a boolean value, but got replacement={self.replacement}")

    @property
    def num_samples(self):
        return len(self.data_source) if self._num_samples is None else self._num_samples

    def __iter__(self):
        n = len(self.data_source)
        g = torch.Generator()
        if self.epoch >= 0:
            g.manual_seed(self.epoch)
        if self.replacement:
            return iter(torch.randint(high=n, size=(self.num_samples,), dtype=torch.int64, generator=g).tolist())
        return iter(torch.randperm(n, generator=g).tolist())

    def __len__(self):
        return self.num_samples

    def set_epoch(self, epoch):
        self.epoch = epoch


class DistributedBatchSampler(data.sampler.BatchSampler):
    def __init__(self, sampler, batch_size, drop_last, rank=-1, world_size=2, wrap_last=False, interleave=False):
        super().__init__(sampler, batch_size, drop_last)
        if rank == -1:
            rank = torch.distributed.get_rank()
        self.rank, self.world_size = rank, world_size
        self.sampler.wrap_around = 0
        self.wrap_around = 0
        self.wrap_last = wrap_last
        self.start_iter = 0
        self.interleave = interleave

    def __iter__(self):
        seq = isinstance(self.sampler, data.SequentialSampler) and self.wrap_around == 0 and not self.wrap_last
        if seq:
            n, bs = len(self.sampler), self.batch_size
            first = self.start_iter
            self.start_iter = 0
            for b in range(first, n // bs):
                yield self._batch(list(range(b * bs, (b + 1) * bs)))
            rest = n % bs
            if rest and not self.drop_last:
                yield self._batch(list(range(n - rest, n)))
            return
        batch, i = [], 0
        for idx in self.data_iterator(self.sampler, wrap_around=False):
            batch.append(idx)
            if len(batch) == self.batch_size:
                tb = self._batch(batch)
                if i >= self.start_iter:
                    yield tb
                    self.start_iter = 0
                i += 1
                batch = []
        if len(batch) > 0 and not self.drop_last:
            if self.wrap_last:
                self.sampler.wrap_around -= self.batch_size
                self.wrap_around += len(batch)
                self.wrap_around %= self.batch_size
            yield self._batch(batch)
        if self.wrap_last:
            self.sampler.wrap_around += self.batch_size

    def data_iterator(self, _iter, wrap_around=False):
        for i, idx in enumerate(_iter):
            if i < self.wrap_around % self.batch_size:
                continue
            if wrap_around:
                self.wrap_around += 1
                self.wrap_around %= self.batch_size
            yield idx

    def _batch(self, batch):
        if self.interleave:
            return batch[self.rank: self.batch_size: self.world_size]
        start = self.rank * self.batch_size // self.world_size
        end = (self.rank + 1) * self.batch_size // self.world_size
        return batch[start:end]
