"""Learning-rate schedules used by ReLoRA.

Behavioural parity target: reference ``peft_pretraining/training_utils.py:56-236``
(``get_scheculer`` [sic], cyclical cosine with a floor, and the "jagged" cosine with
a re-warm-up after every restart).  The schedules are expressed here as small
stateless multiplier objects so that the same object can drive

* a ``torch.optim.lr_scheduler.LambdaLR`` (checkpoint-compatible ``state_dict``), and
* the fused AdamW kernels, which take the learning rate as a device scalar.

Golden values for the multipliers live in ``tests/test_schedulers.py``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

from torch.optim.lr_scheduler import LambdaLR

__all__ = [
    "LinearWarmupDecay",
    "CyclicalCosine",
    "JaggedCosine",
    "build_multiplier",
    "get_scheduler",
    "get_scheculer",
]


def _cosine(progress: float, floor: float) -> float:
    return floor + (1.0 - floor) * 0.5 * (1.0 + math.cos(math.pi * progress))


@dataclass(frozen=True)
class LinearWarmupDecay:
    """Linear warm-up to 1 then linear decay to 0 (HF ``get_linear_schedule_with_warmup``)."""

    warmup: int
    total: int

    def __call__(self, step: int) -> float:
        if step < self.warmup:
            return step / max(1, self.warmup)
        return max(0.0, (self.total - step) / max(1, self.total - self.warmup))


@dataclass(frozen=True)
class CyclicalCosine:
    """Cosine cycles of ``cycle`` steps, each with ``warmup`` linear steps and a floor.

    From the second cycle on, the first two steps of a cycle use a 1e-7 multiplier
    instead of 0 and 1/warmup (reference ``training_utils.py:179-183``).
    """

    warmup: int
    cycle: int
    floor: float = 0.1

    def __post_init__(self):
        if not 0.0 < self.floor <= 1.0:
            raise ValueError("min_lr_ratio must be in (0,1]")

    def __call__(self, step: int) -> float:
        pos = step % self.cycle
        if pos < self.warmup:
            if step != pos and pos < 2:
                return 1e-7
            return pos / max(1, self.warmup)
        return _cosine((pos - self.warmup) / max(1, self.cycle - self.warmup), self.floor)


@dataclass(frozen=True)
class JaggedCosine:
    """One global cosine decay, re-warmed from 0 after every ``restart_every`` steps.

    ``adjust`` shifts the restart grid (used to line restarts up with a warm-started
    run).  A restart is only re-warmed once ``step >= restart_every`` (so an adjusted
    first restart is *not* warmed — reference ``training_utils.py:221``).
    """

    total: int
    first_warmup: int
    restart_warmup: int
    restart_every: int
    floor: float = 0.1
    adjust: int = 0

    def __post_init__(self):
        if not 0.0 < self.floor <= 1.0:
            raise ValueError("min_lr_ratio must be in (0,1]")
        if self.restart_every <= 0:
            raise ValueError("restart_every must be positive")
        if self.adjust + self.first_warmup > self.total:
            raise ValueError("warmup + adjust_step is more than full training steps")
        if self.adjust + self.first_warmup > self.restart_every:
            raise ValueError("the first reset will happen before the warmup is done")

    def __call__(self, step: int) -> float:
        if step < self.first_warmup:
            return step / max(1, self.first_warmup)
        shifted = step + self.adjust
        pos = shifted % self.restart_every
        span = max(1, self.total - self.first_warmup)
        if pos < self.restart_warmup and step >= self.restart_every:
            k = shifted // self.restart_every
            peak = _cosine((k * self.restart_every + self.restart_warmup - self.first_warmup) / span, self.floor)
            return pos / max(1, self.restart_warmup) * peak
        return _cosine((shifted - self.first_warmup) / span, self.floor)


def build_multiplier(
    scheduler_type: str,
    *,
    num_training_steps: int,
    warmup_steps: int,
    min_lr_ratio: float = 0.1,
    cycle_length: Optional[int] = None,
    restart_warmup_steps: Optional[int] = None,
    adjust_step: int = 0,
):
    """Return the step -> multiplier callable for ``scheduler_type``."""
    if adjust_step != 0 and scheduler_type != "cosine_restarts":
        raise ValueError("adjust_step is only supported for cosine_restarts scheduler")
    if scheduler_type == "linear":
        return LinearWarmupDecay(warmup_steps, num_training_steps)
    if scheduler_type == "cosine":
        cycle = num_training_steps if cycle_length is None else cycle_length
        if num_training_steps % cycle != 0:
            raise ValueError(
                f"num_training_steps ({num_training_steps}) must be divisible by cycle_length ({cycle})"
            )
        return CyclicalCosine(warmup_steps, cycle, min_lr_ratio)
    if scheduler_type == "cosine_restarts":
        if restart_warmup_steps is None:
            raise ValueError("restart_warmup_steps must be specified for cosine_restarts scheduler")
        if cycle_length is None:
            raise ValueError("restart_every must be specified for cosine_restarts scheduler")
        if num_training_steps % cycle_length != 0:
            raise ValueError(
                f"num_training_steps ({num_training_steps}) must be divisible by restart_every ({cycle_length})"
            )
        return JaggedCosine(
            total=num_training_steps,
            first_warmup=warmup_steps,
            restart_warmup=restart_warmup_steps,
            restart_every=cycle_length,
            floor=min_lr_ratio,
            adjust=adjust_step,
        )
    raise NotImplementedError(f"Scheduler {scheduler_type} is not implemented")


def get_scheduler(
    optimizer,
    *,
    scheduler_type: str,
    num_training_steps: int,
    warmup_steps: int,
    min_lr_ratio: float,
    cycle_length: Optional[int] = None,
    restart_warmup_steps: Optional[int] = None,
    adjust_step: int = 0,
    last_epoch: int = -1,
) -> LambdaLR:
    """LambdaLR over :func:`build_multiplier` (same keyword surface as the reference)."""
    mult = build_multiplier(
        scheduler_type,
        num_training_steps=num_training_steps,
        warmup_steps=warmup_steps,
        min_lr_ratio=min_lr_ratio,
        cycle_length=cycle_length,
        restart_warmup_steps=restart_warmup_steps,
        adjust_step=adjust_step,
    )
    def lr_lambda(step: int) -> float:
        # a plain function: LambdaLR.state_dict() serialises the __dict__ of callable *objects* and load_state_dict() would
        # overwrite the freshly built schedule with the checkpoint's (total, warmup, restart_every ...) on resume; like the
        # reference's functools.partial, only last_epoch / base_lrs are restored and new CLI values win
        return mult(step)

    lr_lambda.multiplier = mult  # introspection (tools/plot_lr.py, tests)
    return LambdaLR(optimizer, lr_lambda, last_epoch)


# the reference spells it this way (training_utils.py:56); keep the alias for drop-in use
get_scheculer = get_scheduler
