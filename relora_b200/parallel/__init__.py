"""Single-box data parallelism: process-group plumbing, flat parameter store, gradient sync,
ZeRO-1 optimizer-state sharding and the NVLink peer-memory collectives."""
from .dist import DistInfo, barrier, broadcast_object, init_distributed, shutdown
from .flat import FlatAdamW, FlatParamStore
from .grad_sync import GradSync, broadcast_params

__all__ = [
    "DistInfo",
    "init_distributed",
    "barrier",
    "broadcast_object",
    "shutdown",
    "FlatParamStore",
    "FlatAdamW",
    "GradSync",
    "broadcast_params",
]
