"""Training engine: orchestration (``trainer.run``), steppers, fused Llama executor."""
from .stepper import ModuleStepper, UpdateInfo, make_stepper
from .trainer import TrainState, evaluate_model, run

__all__ = ["run", "evaluate_model", "TrainState", "ModuleStepper", "UpdateInfo", "make_stepper"]
