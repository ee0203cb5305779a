"""Batch-level distributed sampling for the NeoX data path, with O(1) resume.

Behavioural target: ``megatron_dataset/samplers.py`` of the reference (``RandomSampler`` ``:29-86``,
``DistributedBatchSampler`` ``:89-165``).  Every rank walks the same stream of *global* batches and keeps its own slice of each
one (contiguous by default, strided with ``interleave=True``).  ``start_iter`` is the number of global batches already
consumed: with a sequential index stream the skip is pure index arithmetic (the reference builds and throws the batches
away); ``wrap_last`` carries the tail of an epoch into the next one exactly like upstream.
"""
from __future__ import annotations

from typing import Iterator, List, Optional, Sequence

import torch
from torch.utils.data import BatchSampler, Sampler, SequentialSampler

__all__ = ["DistributedBatchSampler", "RandomSampler"]


class RandomSampler(Sampler):
    """Index sampler reseeded per epoch through :meth:`set_epoch`.

    Without replacement an epoch is one permutation of the data source; with replacement ``num_samples`` independent draws.
    Before the first ``set_epoch`` (epoch ``-1``) the generator is left unseeded, as upstream does.
    """

    def __init__(self, data_source: Sequence, replacement: bool = False, num_samples: Optional[int] = None):
        if not isinstance(replacement, bool):
            raise ValueError(f"replacement must be a bool, got {replacement!r}")
        if num_samples is not None and not replacement:
            raise ValueError("num_samples only makes sense together with replacement=True (a permutation has a fixed length)")
        self.data_source = data_source
        self.replacement = replacement
        self._requested = num_samples
        self.epoch = -1
        if not isinstance(self.num_samples, int) or self.num_samples <= 0:
            raise ValueError(f"the sampler needs a positive number of samples, got {self.num_samples!r}")

    @property
    def num_samples(self) -> int:
        return self._requested if self._requested is not None else len(self.data_source)

    def set_epoch(self, epoch: int) -> None:
        self.epoch = epoch

    def _generator(self) -> torch.Generator:
        gen = torch.Generator()
        if self.epoch >= 0:
            gen.manual_seed(self.epoch)
        return gen

    def __iter__(self) -> Iterator[int]:
        population = len(self.data_source)
        gen = self._generator()
        if self.replacement:
            draws = torch.randint(high=population, size=(self.num_samples,), dtype=torch.int64, generator=gen)
        else:
            draws = torch.randperm(population, generator=gen)
        return iter(draws.tolist())

    def __len__(self) -> int:
        return self.num_samples


class DistributedBatchSampler(BatchSampler):
    """Yields, for every global batch of ``batch_size`` indices, the part that belongs to ``rank``."""

    def __init__(self, sampler, batch_size: int, drop_last: bool, rank: int = -1, world_size: int = 2, wrap_last: bool = False,
                 interleave: bool = False):
        super().__init__(sampler, batch_size, drop_last)
        self.rank = torch.distributed.get_rank() if rank == -1 else rank
        self.world_size = world_size
        self.wrap_last = wrap_last
        self.interleave = interleave
        self.start_iter = 0       # global batches to skip at the next __iter__ (resume)
        self.wrap_around = 0      # indices of the current epoch already handed out by a wrapped last batch
        self.sampler.wrap_around = 0

    # ------------------------------------------------------------------ slicing
    def _batch(self, global_batch: List[int]) -> List[int]:
        if self.interleave:
            return global_batch[self.rank:self.batch_size:self.world_size]
        lo = self.rank * self.batch_size // self.world_size
        hi = (self.rank + 1) * self.batch_size // self.world_size
        return global_batch[lo:hi]

    # ------------------------------------------------------------------ index stream
    def data_iterator(self, index_iter, wrap_around: bool = False) -> Iterator[int]:
        """The sampler's indices minus the ones a wrapped batch of the previous epoch already consumed."""
        for position, index in enumerate(index_iter):
            if position < self.wrap_around % self.batch_size:
                continue
            if wrap_around:
                self.wrap_around = (self.wrap_around + 1) % self.batch_size
            yield index

    def _fast_path(self) -> bool:
        return isinstance(self.sampler, SequentialSampler) and self.wrap_around == 0 and not self.wrap_last

    def __iter__(self) -> Iterator[List[int]]:
        if self._fast_path():
            # sequential stream: batch b is range(b*bs, (b+1)*bs), so resuming is an offset
            total, bs = len(self.sampler), self.batch_size
            first, self.start_iter = self.start_iter, 0
            for b in range(first, total // bs):
                yield self._batch(list(range(b * bs, b * bs + bs)))
            tail = total % bs
            if tail and not self.drop_last:
                yield self._batch(list(range(total - tail, total)))
            return

        pending: List[int] = []
        produced = 0
        for index in self.data_iterator(self.sampler, wrap_around=False):
            pending.append(index)
            if len(pending) < self.batch_size:
                continue
            if produced >= self.start_iter:
                yield self._batch(pending)
                self.start_iter = 0
            produced += 1
            pending = []
        if pending and not self.drop_last:
            if self.wrap_last:
                self.sampler.wrap_around -= self.batch_size
                self.wrap_around = (self.wrap_around + len(pending)) % self.batch_size
            yield self._batch(pending)
        if self.wrap_last:
            self.sampler.wrap_around += self.batch_size
