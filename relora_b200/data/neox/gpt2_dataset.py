"""GPT-style sample view over an indexed token store: fixed ``seq_length + 1``-token samples cut from the
shuffled concatenation of documents (reference ``megatron_dataset/dataset.py:32-330``).

Three index maps, cached next to the data as ``<prefix>_<name>_indexmap_<N>ns_<S>sl_<seed>s_{doc,sample,shuffle}_idx.npy``:
``doc_idx`` (epoch-replicated, shuffled document order, numpy ``RandomState(seed)``), ``sample_idx`` (built by the
native helper), ``shuffle_idx`` (sample permutation).  Rank 0 builds, everyone mmap-loads after a barrier on
whatever backend is active (the reference "barriers" with a CUDA all-reduce, ``dataset.py:220-225``).
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch
import torch.distributed as dist

from ...obs import logger
from . import helpers_build

__all__ = ["GPT2Dataset", "build_index_mappings", "num_epochs_for", "build_doc_idx", "build_shuffle_idx", "build_sample_idx_python"]


def num_epochs_for(tokens_per_epoch: int, seq_length: int, num_samples: int) -> int:
    epochs, total = 0, 0
    while True:
        epochs += 1
        total += tokens_per_epoch
        if (total - 1) // seq_length >= num_samples:
            return epochs


def build_doc_idx(documents, num_epochs: int, rng: np.random.RandomState) -> np.ndarray:
    doc_idx = np.tile(np.asarray(documents, dtype=np.int32), num_epochs)
    rng.shuffle(doc_idx)
    return doc_idx


def build_shuffle_idx(size: int, rng: np.random.RandomState) -> np.ndarray:
    dtype = np.uint32 if size < np.iinfo(np.uint32).max - 1 else np.int64
    idx = np.arange(size, dtype=dtype)
    rng.shuffle(idx)
    return idx


def build_sample_idx_python(sizes, doc_idx, seq_length, num_epochs, tokens_per_epoch) -> np.ndarray:
    """Pure-numpy twin of the native ``build_sample_idx`` (oracle for tests, fallback without a compiler)."""
    n = (num_epochs * tokens_per_epoch - 1) // seq_length
    out = np.zeros((n + 1, 2), dtype=np.int64)
    cur, off = 0, 0
    for s in range(1, n + 1):
        need = seq_length + 1
        while need > 0:
            avail = int(sizes[doc_idx[cur]]) - off
            if avail >= need:
                off += need - 1
                need = 0
            else:
                need -= avail
                cur += 1
                off = 0
        out[s] = (cur, off)
    return out


def _barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def build_index_mappings(name, data_prefix, documents, sizes, num_samples, seq_length, seed, use_shared_fs=True):
    tokens_per_epoch = int(np.sum(sizes[documents]))
    num_epochs = num_epochs_for(tokens_per_epoch, seq_length, num_samples)
    rng = np.random.RandomState(seed=seed)
    stem = f"{data_prefix}_{name}_indexmap_{num_samples}ns_{seq_length}sl_{seed}s"
    files = {k: f"{stem}_{k}_idx.npy" for k in ("doc", "sample", "shuffle")}

    if not use_shared_fs:
        builder = int(os.environ.get("LOCAL_RANK", "0")) == 0
    elif dist.is_initialized():
        builder = dist.get_rank() == 0
    else:
        builder = True
    if builder and not all(os.path.isfile(f) for f in files.values()):
        logger.warning(" > WARNING: could not find index map files, building the indices on rank 0 ...")
        t0 = time.time()
        doc_idx = build_doc_idx(documents, num_epochs, rng)
        np.save(files["doc"], doc_idx, allow_pickle=True)
        logger.info(f" > elapsed time to build and save doc-idx mapping (seconds): {time.time() - t0:4f}")
        t0 = time.time()
        sizes32 = np.ascontiguousarray(sizes, dtype=np.int32)
        helpers = helpers_build.load()
        n_samples = (num_epochs * tokens_per_epoch - 1) / seq_length
        fn = helpers.build_sample_idx_int32 if 2 * (n_samples + 1) < np.iinfo(np.int32).max else helpers.build_sample_idx_int64
        sample_idx = fn(sizes32, doc_idx, seq_length, num_epochs, tokens_per_epoch)
        np.save(files["sample"], sample_idx, allow_pickle=True)
        logger.info(f" > elapsed time to build and save sample-idx mapping (seconds): {time.time() - t0:4f}")
        t0 = time.time()
        shuffle_idx = build_shuffle_idx(sample_idx.shape[0] - 1, rng)
        np.save(files["shuffle"], shuffle_idx, allow_pickle=True)
        logger.info(f" > elapsed time to build and save shuffle-idx mapping (seconds): {time.time() - t0:4f}")
    _barrier()
    t0 = time.time()
    doc_idx = np.load(files["doc"], allow_pickle=True, mmap_mode="r")
    sample_idx = np.load(files["sample"], allow_pickle=True, mmap_mode="r")
    shuffle_idx = np.load(files["shuffle"], allow_pickle=True, mmap_mode="r")
    logger.info(f"    loaded indexed file in {time.time() - t0:3.3f} seconds")
    logger.info(f"    total number of samples: {sample_idx.shape[0]}")
    logger.info(f"    total number of epochs: {num_epochs}")
    return doc_idx, sample_idx, shuffle_idx


class GPT2Dataset(torch.utils.data.Dataset):
    def __init__(self, name, data_prefix, documents, indexed_dataset, num_samples, seq_length, seed,
                 build_index_mappings_flag=True, use_shared_fs=True, label_dataset=None, build_index_mappings=None):
        if build_index_mappings is not None:  # reference keyword
            build_index_mappings_flag = build_index_mappings
        self.name = name
        self.indexed_dataset = indexed_dataset
        self.label_dataset = label_dataset
        assert np.min(documents) >= 0
        assert np.max(documents) < indexed_dataset.sizes.shape[0]
        if build_index_mappings_flag:
            self.doc_idx, self.sample_idx, self.shuffle_idx = _build(name, data_prefix, documents, indexed_dataset.sizes,
                                                                     num_samples, seq_length, seed, use_shared_fs)
            self.shuffle_idx_len = self.shuffle_idx.shape[0] - 1
            self.sample_idx_len = self.sample_idx.shape[0] - 1
            if self.shuffle_idx_len != self.sample_idx_len - 1:
                logger.warning(f"shuffle index length ({self.shuffle_idx_len}) is not equal to sample index length ({self.sample_idx_len})")

    def __len__(self):
        return min(self.shuffle_idx_len, self.sample_idx_len)

    def __getitem__(self, idx):
        try:
            return self._get(idx)
        except IndexError:
            new_idx = idx % len(self)
            logger.warning(f"Got index out of bounds error with index {idx} - taking modulo of index instead ({new_idx})")
            return self[new_idx]

    def _get(self, idx):
        idx = self.shuffle_idx[idx]
        d_first, off_first = self.sample_idx[idx]
        d_last, off_last = self.sample_idx[idx + 1]
        sources = [self.indexed_dataset] if self.label_dataset is None else [self.indexed_dataset, self.label_dataset]
        outs = []
        for ds in sources:
            if d_first == d_last:
                outs.append(ds.get(self.doc_idx[d_first], offset=int(off_first), length=int(off_last - off_first + 1)))
            else:
                parts = [ds.get(self.doc_idx[d_first], offset=int(off_first))]
                parts.extend(ds.get(self.doc_idx[i]) for i in range(d_first + 1, d_last))
                parts.append(ds.get(self.doc_idx[d_last], length=int(off_last + 1)))
                outs.append(np.concatenate(parts))
        item = {"input_ids": np.array(outs[0], dtype=np.int64)}
        if len(outs) == 2:
            item["label"] = np.array(outs[1], dtype=np.int64)
        return item


_build = build_index_mappings
