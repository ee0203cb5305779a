"""GPT-NeoX / Megatron data path: mmap token stores, GPT sample index (native builders), blending, batch sampling."""
from __future__ import annotations

import yaml

from ...obs import logger
from .args import NeoXArgs
from .blendable import BlendableDataset
from .data_utils import build_train_valid_test_dataloaders
from .gpt2_dataset import GPT2Dataset
from .indexed_dataset import MMapIndexedDataset, MMapIndexedDatasetBuilder, make_builder, make_dataset
from .samplers import DistributedBatchSampler, RandomSampler

__all__ = ["NeoXArgs", "BlendableDataset", "GPT2Dataset", "MMapIndexedDataset", "MMapIndexedDatasetBuilder", "make_builder",
           "make_dataset", "DistributedBatchSampler", "RandomSampler", "build_train_valid_test_dataloaders", "load_megatron_dataset"]


def load_megatron_dataset(args, world_size: int, rank: int = 0, start_iteration: int = 0):
    """``--megatron_dataset_config`` entry point (reference ``torchrun_main.py:276-319``).

    Returns ``(train_loader, eval_loader, test_loader, tokenizer_name, vocab_size_or_None)``; may overwrite
    ``args.max_length`` with the dataset's ``seq_length`` like the reference does."""
    logger.info(f"Loading Megatron dataset arguments from {args.megatron_dataset_config}")
    with open(args.megatron_dataset_config) as f:
        conf = {str(k).replace("-", "_"): v for k, v in (yaml.safe_load(f) or {}).items()}  # NeoX YAMLs use either spelling
    conf["global_num_gpus"] = world_size
    conf["train_micro_batch_size_per_gpu"] = args.batch_size
    conf["gradient_accumulation_steps"] = args.gradient_accumulation
    conf["train_batch_size"] = args.total_batch_size
    conf["num_workers"] = args.workers
    if args.max_length != conf["seq_length"]:
        logger.warning(f"args.max_length ({args.max_length}) does not match seq_length ({conf['seq_length']}) in the dataset config")
        logger.warning("Overwriting max_length with seq_length")
        args.max_length = conf["seq_length"]
    if args.num_training_steps > conf["train_iters"]:
        logger.error(f"num_training_steps ({args.num_training_steps}) is greater than train_iters ({conf['train_iters']})")
        raise ValueError("num_training_steps must be less than train_iters")
    vocab = None
    vocab_file = conf.get("vocab_file")
    if vocab_file:
        try:
            from tokenizers import Tokenizer

            vocab = Tokenizer.from_file(vocab_file).get_vocab_size()
        except Exception as e:
            logger.warning(f"Could not read tokenizer {vocab_file!r} ({type(e).__name__}); skipping the vocab-size check")
    logger.info("*" * 40)
    logger.info("Dataset arguments:")
    for k, v in conf.items():
        logger.info(f"{k:30} {v}")
    logger.info("*" * 40)
    logger.info("Building Megatron dataset")
    neox = NeoXArgs.from_dict(conf)
    if neox.iteration is None:
        neox.iteration = start_iteration
    if neox.train_batch_size != args.total_batch_size:
        raise ValueError("megatron_dataset_args.train_batch_size must match total_batch_size")
    train_loader, eval_loader, test_loader = build_train_valid_test_dataloaders(neox_args=neox)
    logger.info("Megatron dataset built")
    return train_loader, eval_loader, test_loader, vocab_file, vocab
