"""Small run-time helpers of the trainer: optimizer-state accounting, the post-reset learning-rate alarm, memory accounting of
the frozen weights.

``optimizer_state_size`` / ``print_optimizer_state_size`` and ``check_lr_and_alert`` exist upstream
(``peft_pretraining/training_utils.py:367-404``) but are never called there; here the trainer calls the first after every optimizer
reset (debug log) and the second right after the reset, with the alert going to the metrics sink (wandb or JSONL) instead of
``wandb.alert`` only.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from ..obs import logger

__all__ = ["optimizer_state_size", "print_optimizer_state_size", "check_lr_and_alert", "frozen_weight_bytes"]


def optimizer_state_size(optimizer) -> Dict[str, float]:
    """Number of first / second moment entries held by *this rank* and how many of them are non-zero.

    Works for :class:`relora_b200.parallel.flat.FlatAdamW` (flat buffers; a ZeRO-1 rank reports its shard) and for any
    ``torch.optim`` optimizer with ``exp_avg`` / ``exp_avg_sq`` state."""
    m = getattr(optimizer, "exp_avg", None)
    v = getattr(optimizer, "exp_avg_sq", None)
    if torch.is_tensor(m) and torch.is_tensor(v):
        tensors = [(m, v)]
    else:
        tensors = [(st["exp_avg"], st["exp_avg_sq"]) for st in optimizer.state.values() if len(st) > 0 and "exp_avg" in st]
    n1 = sum(int(a.numel()) for a, _ in tensors)
    n2 = sum(int(b.numel()) for _, b in tensors)
    nz1 = sum(int((a != 0).sum()) for a, _ in tensors)
    nz2 = sum(int((b != 0).sum()) for _, b in tensors)
    return {"exp_avg_numel": n1, "exp_avg_sq_numel": n2, "exp_avg_nonzero": nz1, "exp_avg_sq_nonzero": nz2}


def print_optimizer_state_size(optimizer, rank: int = 0) -> Dict[str, float]:
    s = optimizer_state_size(optimizer)
    print(f"(Rank {rank}) Number of floats in the first moment: {s['exp_avg_numel'] / 1_000_000:.2f}M "
          f"({s['exp_avg_nonzero'] / 1_000_000:.2f}M non-zero)")
    print(f"(Rank {rank}) Number of floats in the second moment: {s['exp_avg_sq_numel'] / 1_000_000:.2f}M "
          f"({s['exp_avg_sq_nonzero'] / 1_000_000:.2f}M non-zero)")
    return s


def check_lr_and_alert(optimizer, max_lr: float, sink=None, step: Optional[int] = None) -> bool:
    """True (and a warning + a sink record) when the learning rate right after a reset exceeds ``max_lr``: with a restart warm-up
    the first update after the reset must run at (nearly) zero lr, a large value means schedule and reset are out of phase."""
    lr = float(optimizer.param_groups[0]["lr"])
    if lr <= max_lr:
        return False
    msg = f"Optimizer lr after the reset is large. This can lead to instability. Current lr is {lr}"
    logger.warning(msg)
    if sink is not None:
        try:
            sink.alert("Learning rate issue", msg)
        except Exception:
            sink.log({"alert": msg}, step=step)
    return True


def frozen_weight_bytes(model: torch.nn.Module) -> Dict[str, int]:
    """Resident bytes of the frozen ReLoRA weights (packed size for block-scaled storage) vs their bf16 size."""
    from ..relora.linear import ReLoRaLinear

    resident = dense_bf16 = 0
    for m in model.modules():
        if isinstance(m, ReLoRaLinear) and not m.lora_only:
            resident += m.frozen_weight_nbytes()
            dense_bf16 += 2 * m.in_features * m.out_features
    return {"resident_bytes": resident, "bf16_bytes": dense_bf16}
