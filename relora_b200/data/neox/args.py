"""``NeoXArgs`` — the slice of the GPT-NeoX argument system the data path consumes.

The reference vendors ~2.8 K lines of NeoX dataclasses (``megatron_dataset/arguments.py``, ``neox_args.py``) of
which the loaders read about twenty-five fields.  This class keeps the same construction surface
(``NeoXArgs.from_dict`` / ``from_ymls``), the same derived batch arithmetic (``calculate_batch_parameters``,
``arguments.py:753-791``) and the same validation of those fields; every other key of a NeoX YAML is accepted
and stored verbatim (model / optimizer / DeepSpeed sections are irrelevant to data loading), so existing
``pile_megatron_dataset.yaml``-style files work unchanged — including ones without the dummy model settings the
upstream validator insists on (``arguments.py:1077-1106``).
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import yaml

__all__ = ["NeoXArgs"]

_DEFAULTS: Dict[str, Any] = dict(
    # data
    data_path=None, train_data_paths=None, valid_data_paths=None, test_data_paths=None, label_data_paths=None,
    train_data_weights=None, valid_data_weights=None, test_data_weights=None, weight_by_num_documents=False,
    weighted_sampler_alpha=0.3, data_impl="infer", mmap_warmup=False, split="969, 30, 1", seq_length=None,
    use_shared_fs=True, vocab_file=None, merge_file=None, tokenizer_type="GPT2BPETokenizer",
    # schedule
    seed=1234, train_iters=None, eval_interval=1000, eval_iters=100, iteration=None, num_workers=2,
    # batch
    train_batch_size=None, train_micro_batch_size_per_gpu=None, gradient_accumulation_steps=None, batch_size=None,
    # topology
    global_num_gpus=None, pipe_parallel_size=0, model_parallel_size=1, is_pipe_parallel=False,
    # bookkeeping set by the loaders
    do_train=None, do_valid=None, do_test=None,
)


class NeoXArgs:
    def __init__(self, **kwargs):
        for k, v in _DEFAULTS.items():
            setattr(self, k, v)
        self._extra_keys: List[str] = []
        for k, v in kwargs.items():
            key = k.replace("-", "_")
            if key not in _DEFAULTS:
                self._extra_keys.append(key)
            setattr(self, key, v)
        self.calculate_derived()
        self.validate()

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_dict(cls, args_dict: Dict[str, Any]) -> "NeoXArgs":
        return cls(**args_dict)

    @classmethod
    def from_ymls(cls, paths: List[str], overwrite_values: Optional[Dict[str, Any]] = None) -> "NeoXArgs":
        merged: Dict[str, Any] = {}
        for p in paths:
            with open(p) as f:
                conf = yaml.safe_load(f) or {}
            for k, v in conf.items():
                key = k.replace("-", "_")
                if key in merged:
                    raise ValueError(f"Conf file {p} has the following duplicate keys with previously loaded file: {key}")
                merged[key] = v
        merged.update(overwrite_values or {})
        return cls(**merged)

    def update_value(self, key: str, value: Any) -> None:
        setattr(self, key, value)

    def update_values(self, d: Dict[str, Any]) -> None:
        for k, v in d.items():
            setattr(self, k, v)

    def all_config(self) -> Dict[str, Any]:
        return {k: v for k, v in vars(self).items() if not k.startswith("_")}

    # ------------------------------------------------------------------ derivation
    @staticmethod
    def calculate_batch_parameters(dp_world_size, train_batch=None, micro_batch=None, grad_acc=None):
        """Fill in whichever of (global batch, micro batch, accumulation) is missing."""
        if train_batch is not None and micro_batch is not None and grad_acc is not None:
            return train_batch, micro_batch, grad_acc
        if train_batch is not None and micro_batch is not None:
            grad_acc = (train_batch // micro_batch) // dp_world_size
        elif train_batch is not None and grad_acc is not None:
            micro_batch = (train_batch // dp_world_size) // grad_acc
        elif micro_batch is not None and grad_acc is not None:
            train_batch = micro_batch * grad_acc * dp_world_size
        elif train_batch is not None:
            grad_acc, micro_batch = 1, train_batch // dp_world_size
        elif micro_batch is not None:
            train_batch, grad_acc = micro_batch * dp_world_size, 1
        else:
            raise AssertionError("Either train_batch_size or train_micro_batch_size_per_gpu needs to be provided")
        return int(train_batch), int(micro_batch), int(grad_acc)

    @staticmethod
    def check_batch_parameters(dp_world_size, train_batch, micro_batch, grad_acc):
        assert train_batch > 0, f"Train batch size: {train_batch} has to be greater than 0"
        assert micro_batch > 0, f"Micro batch size per gpu: {micro_batch} has to be greater than 0"
        assert grad_acc > 0, f"Gradient accumulation steps: {grad_acc} has to be greater than 0"
        assert train_batch == micro_batch * grad_acc * dp_world_size, (
            f"Check batch related parameters. train_batch_size is not equal to micro_batch_per_gpu * gradient_acc_step * "
            f"world_size \n{train_batch} != {micro_batch} * {grad_acc} * {dp_world_size}")

    def calculate_derived(self) -> None:
        if self.global_num_gpus is None:
            raise RuntimeError("ReLoRA dataloading does not support automatic setting of global_num_gpus "
                               "to reduce the amount of megatron dependencies")
        pp = self.pipe_parallel_size if (self.pipe_parallel_size or 0) >= 1 else 1
        mp = self.model_parallel_size if (self.model_parallel_size or 0) >= 1 else 1
        self.model_parallel_size = mp
        dp = (self.global_num_gpus / pp) / mp
        if dp % 1 != 0:
            raise AssertionError(f"(global_num_gpus / pp_size) / mp_size [({self.global_num_gpus} / {pp}) / {mp}] must be a whole number")
        dp = int(dp)
        micro = self.train_micro_batch_size_per_gpu if self.train_micro_batch_size_per_gpu is not None else self.batch_size
        tb, mb, ga = self.calculate_batch_parameters(dp, self.train_batch_size, micro, self.gradient_accumulation_steps)
        self.check_batch_parameters(dp, tb, mb, ga)
        self.train_batch_size, self.train_micro_batch_size_per_gpu, self.gradient_accumulation_steps = tb, mb, ga
        self.batch_size = mb
        self.is_pipe_parallel = (self.pipe_parallel_size or 0) >= 1 and False  # pipeline loading was removed upstream
        for name in ("train", "valid", "test"):
            paths = getattr(self, f"{name}_data_paths")
            if paths is not None and getattr(self, f"{name}_data_weights") is None:
                setattr(self, f"{name}_data_weights", [1.0] * len(paths))

    def validate(self) -> None:
        has_split_paths = all(getattr(self, f"{n}_data_paths") is not None for n in ("train", "valid", "test"))
        if self.data_path is None and not has_split_paths:
            raise ValueError("One of data_path or (train_data_paths, valid_data_paths, test_data_paths) must be provided")
        if self.data_path is not None and has_split_paths:
            raise ValueError("Either data_path or train/valid/test_data_path can be provided, not both")
        if self.seq_length is None:
            raise ValueError("seq_length must be set")
        if self.train_iters is None:
            raise ValueError("train_iters must be set")
        for name in ("train", "valid", "test"):
            paths, weights = getattr(self, f"{name}_data_paths"), getattr(self, f"{name}_data_weights")
            if paths is not None and weights is not None and len(paths) != len(weights):
                raise ValueError(f"{name}_data_weights must have the same length as {name}_data_paths")
        if self.label_data_paths is not None and self.train_data_paths is not None:
            if len(self.label_data_paths) != len(self.train_data_paths):
                raise ValueError("label_data_paths must have the same length as train_data_paths")
