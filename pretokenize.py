"""Pre-tokenise a text dataset into fixed-length blocks for ``torchrun_main.py --dataset_path``.

    python pretokenize.py --tokenizer t5-base --dataset c4 --dataset_config en --text_field text \
        --sequence_length 512 --save_dir preprocessed_data [--take 1000]

CLI and output layout follow the reference ``pretokenize.py`` (``:19-89``): the result is written to
``<save_dir>/<dataset>[_<config>]_<tokenizer>_<sequence_length>`` (refusing to overwrite) together with an ``args.json`` the
trainer validates ``--max_length`` against.  Extras: ``--dataset`` may also be a dataset directory saved with
``save_to_disk`` or a ``.txt`` / ``.jsonl`` file (no network needed), and ``args.json`` records the vocabulary size.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing
import os
import time
from typing import Optional, Sequence

from relora_b200.data import tokenize_and_chunk
from relora_b200.obs import logger

# (flag, type, default, required, help)
_FLAGS = (
    ("--tokenizer", str, None, True, "Hugging Face tokenizer name or local path"),
    ("--dataset", str, None, True, "Hugging Face dataset name, saved dataset directory, or .txt / .jsonl file"),
    ("--dataset_config", str, None, False, "dataset configuration, e.g. 'en' for c4"),
    ("--text_field", str, "text", False, "column holding the raw text"),
    ("--sequence_length", int, 2048, False, "tokens per training sequence"),
    ("--num_cpu", int, multiprocessing.cpu_count(), False, "worker processes of datasets.map"),
    ("--save_dir", str, None, True, "parent directory of the pre-tokenised dataset"),
    ("--take", int, None, False, "keep only the first N examples of every split (the hub dataset is streamed)"),
)


def parse_args(argv: Optional[Sequence[str]] = None) -> argparse.Namespace:
    parser = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    for flag, typ, default, required, text in _FLAGS:
        parser.add_argument(flag, type=typ, default=default, required=required, help=text)
    return parser.parse_args(argv)


def output_dir(args: argparse.Namespace) -> str:
    """``<save_dir>/<dataset>[_<config>]_<tokenizer>_<sequence_length>`` with path separators flattened."""
    source = args.dataset.rstrip("/")
    name = os.path.basename(source) if os.path.exists(source) else args.dataset
    pieces = [name]
    if args.dataset_config is not None:
        pieces.append(args.dataset_config)
    pieces += [args.tokenizer.rstrip("/").replace("/", "_"), str(args.sequence_length)]
    return os.path.join(args.save_dir, "_".join(pieces))


def load_source(args: argparse.Namespace):
    """A ``DatasetDict`` (or streaming dict) from the hub, a saved dataset directory or a plain text / json-lines file."""
    import datasets

    source = args.dataset
    if os.path.isfile(source):
        builder = "json" if source.endswith((".json", ".jsonl")) else "text"
        return datasets.load_dataset(builder, data_files={"train": source})
    if os.path.isdir(source):
        try:
            return datasets.load_from_disk(source)
        except Exception:  # a directory in hub layout rather than save_to_disk layout
            return datasets.load_dataset(source, args.dataset_config)
    return datasets.load_dataset(source, args.dataset_config, streaming=args.take is not None)


def first_n(dataset, n: int):
    """The first ``n`` examples of every split, materialised (streaming splits are drained through a generator)."""
    import datasets

    def cut(split):
        if hasattr(split, "select"):
            return split.select(range(min(n, len(split))))
        return datasets.Dataset.from_generator(lambda: (yield from split.take(n)))

    return datasets.DatasetDict({name: cut(split) for name, split in dataset.items()})


def main(args: argparse.Namespace) -> str:
    from transformers import AutoTokenizer

    logger.info("pretokenize.py configuration:")
    for key, value in sorted(vars(args).items()):
        logger.info(f"    {key} = {value}")
    target = output_dir(args)
    if os.path.exists(target):
        raise ValueError(f"Path {target} already exists")

    tokenizer = AutoTokenizer.from_pretrained(args.tokenizer)
    dataset = load_source(args)
    if args.take is not None:
        logger.info(f"keeping the first {args.take} examples of each split")
        dataset = first_n(dataset, args.take)

    started = time.time()
    blocks = tokenize_and_chunk(tokenizer=tokenizer, dataset=dataset, text_field=args.text_field,
                                sequence_length=args.sequence_length, num_cpu=args.num_cpu)
    logger.info(f"tokenised and chunked in {(time.time() - started) / 60:.1f} min")
    blocks.save_to_disk(target)
    record = dict(vars(args), vocab_size=len(tokenizer))
    with open(os.path.join(target, "args.json"), "w") as fh:
        json.dump(record, fh, indent=4)
    logger.info(f"wrote {target}")
    return target


if __name__ == "__main__":
    main(parse_args())
