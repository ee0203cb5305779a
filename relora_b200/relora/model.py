"""``ReLoRaModel`` — wraps a causal LM, replacing attention/MLP linears with :class:`ReLoRaLinear`.

Parity target: reference ``peft_pretraining/relora.py:18-177``:

* every ``nn.Linear`` whose qualified name contains one of ``target_modules`` is replaced
  (``["attn", "attention", "mlp"]`` in the trainer) — embeddings, norms and the LM head stay
  as ordinary trainable parameters;
* with ``keep_original_weights`` the wrapped network must equal the original at initialisation,
  so ``lora_A`` is zeroed as well (``lora_B`` is already zero) — i.e. *both* factors are zero and
  the low-rank branch receives no gradient until the first ``merge_and_reinit``
  (SURVEY §7.4 item 10).  ``init_lora_a="kaiming"`` opts out of that quirk;
* ``forward`` is the wrapped model's forward; ``save_pretrained`` writes the wrapped model in HF
  layout plus ``relora_config.json``; ``from_pretrained`` rebuilds from those files, including
  the legacy ``keep_original`` key shim.
"""
from __future__ import annotations

import json
import os
from dataclasses import asdict, dataclass
from typing import List, Optional, Union

import torch
import torch.nn as nn

from .linear import ReLoRaLinear

__all__ = ["ReLoRaConfig", "ReLoRaModel", "merge_and_reinit_functional"]


@dataclass
class ReLoRaConfig:
    r: int
    lora_alpha: float
    lora_dropout: float
    target_modules: List[str]
    keep_original_weights: bool
    lora_only: bool = False
    trainable_scaling: bool = False
    quantize: Optional[str] = None
    use_double_quant: bool = False


def merge_and_reinit_functional(module: nn.Module) -> None:
    """``model.apply``-style merge (reference ``relora.py:31-46``; used under FSDP upstream)."""
    if not isinstance(module, ReLoRaLinear):
        return
    if module.quantize is not None:
        raise NotImplementedError(
            "merge_and_reinit_functional for quantized models is not implemented. "
            "Use ReLoRaModel.merge_and_reinit"
        )
    module.merge_and_reinit()


class ReLoRaModel(nn.Module):
    def __init__(
        self,
        model: nn.Module,
        *,
        target_modules: Union[str, List[str]],
        r: int = 128,
        lora_alpha: float = 32,
        lora_dropout: float = 0.1,
        keep_original_weights: bool = True,
        lora_only: bool = False,
        trainable_scaling: bool = False,
        quantize: Optional[str] = None,
        use_double_quant: bool = False,
        init_lora_a: str = "zeros",
    ):
        if r <= 0:
            raise ValueError("r must be positive. If you want r == 0, use the original model.")
        super().__init__()
        if lora_only and keep_original_weights:
            # upstream asserts here (relora.py:126-127), which makes plain `--use_peft` without --relora crash;
            # LoRA-only training has no frozen weight to keep, so the flag is simply dropped
            keep_original_weights = False
        self.wrapped_model = model
        self.r = r
        self.lora_alpha = lora_alpha
        self.lora_dropout = lora_dropout
        self.target_modules = target_modules
        self.keep_original_weights = keep_original_weights
        self.lora_only = lora_only
        self.trainable_scaling = trainable_scaling
        self.n_restarts = 0
        self.seed = 0
        # note: like the reference (relora.py:77-85) lora_only / trainable_scaling are NOT recorded
        self._config = ReLoRaConfig(
            r=r,
            lora_alpha=lora_alpha,
            lora_dropout=lora_dropout,
            target_modules=target_modules,
            keep_original_weights=keep_original_weights,
            quantize=quantize,
            use_double_quant=use_double_quant,
        )

        keys = [target_modules] if isinstance(target_modules, str) else list(target_modules)
        targets = [
            (name, mod)
            for name, mod in self.wrapped_model.named_modules()
            if isinstance(mod, nn.Linear) and any(k in name for k in keys)
        ]
        for index, (name, mod) in enumerate(targets):
            has_bias = mod.bias is not None
            new = ReLoRaLinear(
                mod.in_features,
                mod.out_features,
                bias=has_bias,
                r=r,
                lora_alpha=lora_alpha,
                lora_dropout=lora_dropout,
                lora_only=lora_only,
                trainable_scaling=trainable_scaling,
                quantize=quantize,
                weight_data=mod.weight.data if keep_original_weights else None,
                bias_data=mod.bias.data if (has_bias and keep_original_weights) else None,
                bnb_4bit_use_double_quant=use_double_quant,
                device=mod.weight.device,
                dtype=mod.weight.dtype,
            )
            new.module_index = index
            if keep_original_weights and init_lora_a == "zeros":
                nn.init.zeros_(new.lora_A.weight)
            parent_name, _, leaf = name.rpartition(".")
            parent = self.wrapped_model.get_submodule(parent_name) if parent_name else self.wrapped_model
            setattr(parent, leaf, new)
        self.n_wrapped = len(targets)

    # the reference patches ``self.forward = wrapped.forward``; delegating keeps hooks working
    def forward(self, *args, **kwargs):
        return self.wrapped_model(*args, **kwargs)

    @property
    def config(self):
        return getattr(self.wrapped_model, "config", None)

    def relora_modules(self):
        return [m for m in self.modules() if isinstance(m, ReLoRaLinear)]

    @torch.no_grad()
    def merge_and_reinit(self):
        """Merge every low-rank pair into its frozen weight and restart the factors.

        On CUDA with the fused engine attached this is one batched kernel sequence
        (:func:`relora_b200.ops.dispatch.merge_all`); otherwise a module loop.
        """
        from ..ops import dispatch

        mods = self.relora_modules()
        if not dispatch.merge_all(mods, seed=self.seed, restart_index=self.n_restarts):
            for m in mods:
                m.merge_and_reinit(seed=self.seed, restart_index=self.n_restarts)
        self.n_restarts += 1

    # ------------------------------------------------------------------ persistence
    def save_pretrained(self, path: str, **kwargs):
        self.wrapped_model.save_pretrained(path, **kwargs)
        with open(os.path.join(path, "relora_config.json"), "w") as f:
            json.dump(asdict(self._config), f, indent=4)

    @classmethod
    def from_pretrained(cls, path: str):
        from ..models import model_from_config_dir

        with open(os.path.join(path, "relora_config.json")) as f:
            cfg = json.load(f)
        if "keep_original" in cfg:  # legacy key
            print("WARNING: keep_original is deprecated. Use lora_only instead.")
            cfg["lora_only"] = not cfg.pop("keep_original")
            cfg["keep_original_weights"] = not cfg["lora_only"]
        cfg.setdefault("trainable_scaling", False)
        base = model_from_config_dir(path)
        model = cls(base, **cfg)
        state = torch.load(os.path.join(path, "pytorch_model.bin"), map_location="cpu", weights_only=True)
        model.wrapped_model.load_state_dict(state, strict=True)
        return model
