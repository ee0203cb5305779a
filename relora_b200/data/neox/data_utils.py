"""Dataset / dataloader assembly for the NeoX data path (reference ``megatron_dataset/data_utils.py:34-467``)."""
from __future__ import annotations

import math
from functools import partial
from itertools import zip_longest
from typing import List, Tuple

import numpy as np
import torch
import torch.distributed as dist

from ...obs import logger
from .blendable import BlendableDataset
from .gpt2_dataset import GPT2Dataset
from .indexed_dataset import make_dataset as make_indexed_dataset
from .samplers import DistributedBatchSampler

__all__ = ["make_data_loader", "build_the_dataset", "build_train_valid_test_datasets", "get_train_valid_test_split_",
           "get_normalized_weights_and_num_samples", "build_weighted_datasets", "weights_by_num_docs",
           "build_train_valid_test_dataloaders"]


def make_data_loader(dataset, neox_args):
    """Sequential global batches, each rank takes its contiguous slice; pinned host memory."""
    if dataset is None:
        return None
    world, rank = 1, 0
    if dist.is_initialized():
        world, rank = dist.get_world_size(), dist.get_rank()
    else:
        logger.warning("Not using distributed mode. Should only be used for debugging.")
    sampler = torch.utils.data.SequentialSampler(dataset)
    batch_sampler = DistributedBatchSampler(sampler=sampler, batch_size=neox_args.batch_size * world, drop_last=True,
                                            rank=rank, world_size=world)
    return torch.utils.data.DataLoader(dataset, batch_sampler=batch_sampler, num_workers=neox_args.num_workers,
                                       pin_memory=torch.cuda.is_available())


def build_the_dataset(data_prefix, name, data_impl, num_samples, seq_length, seed, skip_warmup, build_index_mappings=True,
                      label_prefix=None):
    indexed = make_indexed_dataset(data_prefix, data_impl, skip_warmup)
    labels = None if label_prefix is None else make_indexed_dataset(label_prefix, data_impl, skip_warmup)
    n_docs = indexed.sizes.shape[0]
    logger.info(f"    {name}:")
    logger.info(f"     no. of documents:{n_docs}")
    documents = np.arange(n_docs, dtype=np.int32)
    return GPT2Dataset(name, data_prefix, documents, indexed, num_samples, seq_length, seed,
                       build_index_mappings=build_index_mappings, label_dataset=labels)


def get_train_valid_test_split_(splits_string: str, size: int) -> List[int]:
    """'969,30,1' or '969/30/1' → cumulative document boundaries [0, a, b, size]."""
    sep = "," if "," in splits_string else ("/" if "/" in splits_string else None)
    splits = [float(s) for s in splits_string.split(sep)] if sep else [float(splits_string)]
    splits = (splits + [0.0, 0.0, 0.0])[:3]
    total = sum(splits)
    assert total > 0.0
    bounds = [0]
    for s in splits:
        bounds.append(bounds[-1] + int(round(s / total * float(size))))
    diff = bounds[-1] - size
    bounds = [bounds[0]] + [b - diff for b in bounds[1:]]
    assert len(bounds) == 4 and bounds[-1] == size
    return bounds


def build_train_valid_test_datasets(data_prefix, use_shared_fs, data_impl, splits_string, train_valid_test_num_samples,
                                    seq_length, seed, skip_warmup):
    indexed = make_indexed_dataset(data_prefix, data_impl, skip_warmup)
    n_docs = indexed.sizes.shape[0]
    splits = get_train_valid_test_split_(splits_string, n_docs)
    logger.info(" > dataset split:")
    for i, nm in enumerate(("train", "validation", "test")):
        logger.info(f"    {nm}:")
        logger.info(f"     document indices in [{splits[i]}, {splits[i + 1]}) total of {splits[i + 1] - splits[i]} documents")

    def build(i, name):
        if splits[i + 1] <= splits[i]:
            return None
        docs = np.arange(splits[i], splits[i + 1], dtype=np.int32)
        return GPT2Dataset(name, data_prefix, docs, indexed, train_valid_test_num_samples[i], seq_length, seed,
                           use_shared_fs=use_shared_fs)

    return build(0, "train"), build(1, "valid"), build(2, "test")


def get_normalized_weights_and_num_samples(weights: List[float], num_samples: int) -> Tuple[List[float], List[int]]:
    total = sum(weights)
    assert total > 0.0
    weights = [w / total for w in weights]
    # 0.5 % head-room so that blending never runs a component dataset dry
    return weights, [int(math.ceil(num_samples * w * 1.005)) for w in weights]


def build_weighted_datasets(neox_args, train_num_samples, valid_num_samples, test_num_samples, train_weights, valid_weights,
                            test_weights, build_index_mappings=True):
    train, valid, test = [], [], []
    for i, (tr, lab, va, te) in enumerate(zip_longest(neox_args.train_data_paths, neox_args.label_data_paths or [],
                                                      neox_args.valid_data_paths, neox_args.test_data_paths)):
        common = dict(data_impl=neox_args.data_impl, seq_length=neox_args.seq_length, seed=neox_args.seed,
                      skip_warmup=(not neox_args.mmap_warmup), build_index_mappings=build_index_mappings)
        if tr:
            train.append(build_the_dataset(data_prefix=tr, name=f"train_{i}", num_samples=train_num_samples[i], label_prefix=lab, **common))
        if va:
            valid.append(build_the_dataset(data_prefix=va, name=f"valid_{i}", num_samples=valid_num_samples[i], **common))
        if te:
            test.append(build_the_dataset(data_prefix=te, name=f"test_{i}", num_samples=test_num_samples[i], **common))
    return train, valid, test


def weights_by_num_docs(counts: list, alpha: float = 0.3) -> List[float]:
    """p(L) ∝ |L|^α re-weighted by (1 - |L|/Σ) — up-samples small corpora (arXiv:1911.02116)."""
    if len(counts) == 1:
        return [1.0]
    total = sum(counts)
    unbiased = [c / total for c in counts]
    probs = [p ** alpha for p in unbiased]
    z = sum(probs)
    probs = [p / z for p in probs]
    w = [p * (1 - u) for p, u in zip(probs, unbiased)]
    z = sum(w)
    return [x / z for x in w]


def build_train_valid_test_dataloaders(neox_args):
    logger.info("> building train, validation, and test datasets ...")
    assert not neox_args.is_pipe_parallel, "pipeline parallelism was removed from the ReLoRA version of megatron dataloading"
    train_iters = neox_args.train_iters
    eval_iters = (train_iters // neox_args.eval_interval + 1) * neox_args.eval_iters
    test_iters = neox_args.eval_iters
    n_samples = [train_iters * neox_args.train_batch_size, eval_iters * neox_args.train_batch_size, test_iters * neox_args.train_batch_size]
    train_ds = valid_ds = test_ds = None
    if neox_args.train_data_paths:
        tw, tn = get_normalized_weights_and_num_samples(neox_args.train_data_weights, n_samples[0])
        vw, vn = get_normalized_weights_and_num_samples(neox_args.valid_data_weights, n_samples[1])
        sw, sn = get_normalized_weights_and_num_samples(neox_args.test_data_weights, n_samples[2])
        train_sets, valid_sets, test_sets = build_weighted_datasets(neox_args, tn, vn, sn, tw, vw, sw,
                                                                    build_index_mappings=not neox_args.weight_by_num_documents)
        if neox_args.weight_by_num_documents:
            docs = lambda sets: [d.indexed_dataset.sizes.shape[0] for d in sets]  # noqa: E731
            fn = partial(weights_by_num_docs, alpha=neox_args.weighted_sampler_alpha)
            tw, tn = get_normalized_weights_and_num_samples(fn(docs(train_sets)), n_samples[0])
            vw, vn = get_normalized_weights_and_num_samples(fn(docs(valid_sets)), n_samples[1])
            sw, sn = get_normalized_weights_and_num_samples(fn(docs(test_sets)), n_samples[2])
            train_sets, valid_sets, test_sets = build_weighted_datasets(neox_args, tn, vn, sn, tw, vw, sw)
        if train_sets:
            train_ds = BlendableDataset(train_sets, tw)
        if valid_sets:
            valid_ds = BlendableDataset(valid_sets, vw)
        if test_sets:
            test_ds = BlendableDataset(test_sets, sw)
    else:
        train_ds, valid_ds, test_ds = build_train_valid_test_datasets(
            data_prefix=neox_args.data_path, use_shared_fs=neox_args.use_shared_fs, data_impl=neox_args.data_impl,
            splits_string=neox_args.split, train_valid_test_num_samples=n_samples, seq_length=neox_args.seq_length,
            seed=neox_args.seed, skip_warmup=(not neox_args.mmap_warmup))
    train_loader = make_data_loader(train_ds, neox_args)
    valid_loader = make_data_loader(valid_ds, neox_args)
    test_loader = make_data_loader(test_ds, neox_args)
    neox_args.do_train = int(train_loader is not None and neox_args.train_iters > 0)
    neox_args.do_valid = int(valid_loader is not None and neox_args.eval_iters > 0)
    neox_args.do_test = int(test_loader is not None and neox_args.eval_iters > 0)
    it = neox_args.iteration or 0
    if train_loader is not None:
        train_loader.batch_sampler.start_iter = (it * neox_args.gradient_accumulation_steps) % len(train_loader)
        logger.info(f"setting training data start iteration to {train_loader.batch_sampler.start_iter}")
    if valid_loader is not None:
        start = ((it * neox_args.gradient_accumulation_steps) // neox_args.eval_interval) * neox_args.eval_iters
        valid_loader.batch_sampler.start_iter = start % len(valid_loader)
        logger.info(f"setting validation data start iteration to {valid_loader.batch_sampler.start_iter}")
    return train_loader, valid_loader, test_loader
