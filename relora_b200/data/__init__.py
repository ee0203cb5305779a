"""Data pipelines: pre-tokenised HF datasets, synthetic tokens, NeoX/Megatron mmap datasets."""
from .hf_disk import (
    PreprocessedIterableDataset,
    SkipBatchSampler,
    SkipDataLoader,
    check_dataset_size,
    collate_input_ids,
    load_pretokenized,
    shard_for_rank,
    tokenize_and_chunk,
)
from .synthetic import SyntheticTokens, write_synthetic_hf_dataset

__all__ = [
    "PreprocessedIterableDataset",
    "SkipBatchSampler",
    "SkipDataLoader",
    "check_dataset_size",
    "collate_input_ids",
    "load_pretokenized",
    "shard_for_rank",
    "tokenize_and_chunk",
    "SyntheticTokens",
    "write_synthetic_hf_dataset",
]
