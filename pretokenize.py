"""Pre-tokenise a Hugging Face text dataset into fixed-length blocks for ``torchrun_main.py --dataset_path``.

    python pretokenize.py --tokenizer t5-base --dataset c4 --dataset_config en --text_field text \
        --sequence_length 512 --save_dir preprocessed_data [--take 1000]

Same CLI and output layout as the reference ``pretokenize.py``: the dataset is saved under
``<save_dir>/<dataset>[_<config>]_<tokenizer>_<sequence_length>`` (an existing directory is an error) next to an
``args.json`` that the trainer checks its ``--max_length`` against.  Additions: ``--dataset`` may be a local
directory / text file (no network needed), and the tokenizer's vocabulary size is recorded in ``args.json``.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing
import os
import time

from relora_b200.data import tokenize_and_chunk
from relora_b200.obs import logger


def parse_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--tokenizer", type=str, required=True, help="HuggingFace tokenizer name or local path")
    p.add_argument("--dataset", type=str, required=True, help="HuggingFace dataset name, local dataset dir, or .txt/.jsonl file")
    p.add_argument("--dataset_config", type=str, default=None, help="dataset config name, e.g. wikitext-2-v1")
    p.add_argument("--text_field", type=str, default="text")
    p.add_argument("--sequence_length", type=int, default=2048)
    p.add_argument("--num_cpu", type=int, default=multiprocessing.cpu_count())
    p.add_argument("--save_dir", type=str, required=True)
    p.add_argument("--take", type=int, default=None, help="only the first N examples of every split (streams the dataset)")
    return p.parse_args(argv)


def _load(args):
    import datasets

    src = args.dataset
    if os.path.isdir(src):
        try:
            return datasets.load_from_disk(src)
        except Exception:
            return datasets.load_dataset(src, args.dataset_config)
    if os.path.isfile(src):
        kind = "json" if src.endswith((".json", ".jsonl")) else "text"
        return datasets.load_dataset(kind, data_files={"train": src})
    return datasets.load_dataset(src, args.dataset_config, streaming=args.take is not None)


def main(args):
    logger.info("*" * 40)
    logger.info("Starting script with the arguments")
    for k, v in vars(args).items():
        logger.info(f"{k:30} {v}")
    logger.info("*" * 40)
    import datasets
    from transformers import AutoTokenizer

    tok_name = args.tokenizer.rstrip("/").replace("/", "_")
    ds_name = os.path.basename(args.dataset.rstrip("/")) if os.path.exists(args.dataset) else args.dataset
    parts = [ds_name] + ([args.dataset_config] if args.dataset_config is not None else []) + [tok_name, str(args.sequence_length)]
    save_path = os.path.join(args.save_dir, "_".join(parts))
    if os.path.exists(save_path):
        raise ValueError(f"Path {save_path} already exists")

    tokenizer = AutoTokenizer.from_pretrained(args.tokenizer)
    dataset = _load(args)
    if args.take is not None:
        logger.info(f"Taking {args.take} examples from the dataset")

        def head(split):
            if hasattr(split, "take") and not hasattr(split, "select"):
                return datasets.Dataset.from_generator(lambda: (yield from split.take(args.take)))
            return split.select(range(min(args.take, len(split))))

        dataset = datasets.DatasetDict({k: head(v) for k, v in dataset.items()})

    logger.info("Tokenizing and chunking the dataset")
    t0 = time.time()
    dataset = tokenize_and_chunk(tokenizer=tokenizer, dataset=dataset, text_field=args.text_field,
                                 sequence_length=args.sequence_length, num_cpu=args.num_cpu)
    logger.info(f"Tokenization and chunking took {(time.time() - t0) / 3600:.2f} hours")
    dataset.save_to_disk(save_path)
    logger.info(f"Saved the dataset to {save_path}")
    meta = dict(vars(args))
    meta["vocab_size"] = len(tokenizer)
    with open(os.path.join(save_path, "args.json"), "w") as f:
        json.dump(meta, f, indent=4)
    return save_path


if __name__ == "__main__":
    main(parse_args())
