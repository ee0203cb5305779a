from .linear import ReLoRaLinear
from .model import ReLoRaConfig, ReLoRaModel, merge_and_reinit_functional
from .optim_reset import magnitude_pruning_, optimizer_reset, random_pruning_
from .schedulers import build_multiplier, get_scheculer, get_scheduler

__all__ = [
    "ReLoRaLinear",
    "ReLoRaConfig",
    "ReLoRaModel",
    "merge_and_reinit_functional",
    "optimizer_reset",
    "random_pruning_",
    "magnitude_pruning_",
    "get_scheduler",
    "get_scheculer",
    "build_multiplier",
]
