"""Process-group plumbing: one process per GPU, ``torch.distributed`` over NCCL (or gloo on CPU).

Parity target: reference ``torchrun_main.py:344-352`` (env contract, NCCL init), ``:401, :414``
(barriers), ``:416-420`` (run-name broadcast).  Unlike upstream the backend is not hard-coded:
``gloo`` makes the whole trainer runnable on a CPU box (BASELINE.json config 1), and a missing
torchrun environment degrades to a single-process group so scripts and tests can call
``engine.run`` directly.
"""
from __future__ import annotations

import datetime
import os
from dataclasses import dataclass
from typing import Any, List, Optional

import torch
import torch.distributed as dist

__all__ = ["DistInfo", "init_distributed", "barrier", "broadcast_object", "all_reduce_sum_", "shutdown", "is_main"]


@dataclass
class DistInfo:
    rank: int
    local_rank: int
    world_size: int
    device: torch.device
    backend: str
    owns_group: bool = False

    @property
    def is_main(self) -> bool:
        return self.rank == 0


def _pick_device(requested: str, local_rank: int) -> torch.device:
    if requested == "cpu":
        return torch.device("cpu")
    if requested == "cuda" or (requested == "auto" and torch.cuda.is_available()):
        if not torch.cuda.is_available():
            raise RuntimeError("--device cuda requested but no CUDA device is visible")
        torch.cuda.set_device(local_rank)
        return torch.device("cuda", local_rank)
    return torch.device("cpu")


def init_distributed(device: str = "auto", backend: str = "auto", timeout_s: int = 1800) -> DistInfo:
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dev = _pick_device(device, local_rank)
    if backend == "auto":
        backend = "nccl" if dev.type == "cuda" else "gloo"
    owns = False
    if not dist.is_initialized():
        if "MASTER_ADDR" not in os.environ:
            os.environ["MASTER_ADDR"] = "127.0.0.1"
        if "MASTER_PORT" not in os.environ:
            os.environ["MASTER_PORT"] = str(29500 + (os.getpid() % 2000))
        kw = {}
        if backend == "nccl":
            kw["device_id"] = dev
        dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                timeout=datetime.timedelta(seconds=timeout_s), **kw)
        owns = True
    else:
        backend = dist.get_backend()
        rank, world = dist.get_rank(), dist.get_world_size()
    return DistInfo(rank, local_rank, world, dev, backend, owns)


def is_main() -> bool:
    return not dist.is_initialized() or dist.get_rank() == 0


def barrier() -> None:
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def broadcast_object(obj: Any, src: int = 0) -> Any:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return obj
    box: List[Any] = [obj if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def all_reduce_sum_(t: torch.Tensor) -> torch.Tensor:
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def shutdown(info: Optional[DistInfo] = None) -> None:
    if dist.is_initialized() and (info is None or info.owns_group):
        try:
            dist.destroy_process_group()
        except Exception:  # pragma: no cover
            pass
