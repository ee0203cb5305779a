"""Observability: rank-0 logging, metric sinks (wandb / JSONL), torch profiler + NVTX ranges.

Parity target: reference ``torchrun_main.py:322-335`` (profiler), ``:371`` (rank-0 logger),
``:404-412`` (wandb init), ``:923-942`` (per-update metrics).  wandb is optional here: when it
is missing, disabled (``WANDB_MODE=disabled``) or fails to initialise, metrics go to a JSONL
file under the save directory so that offline boxes still get a complete record.
"""
from __future__ import annotations

import contextlib
import json
import os
import sys
import time
import uuid
from typing import Any, Dict, Optional

try:  # loguru is what the reference uses; fall back to stdlib logging if absent
    from loguru import logger as _loguru_logger

    logger = _loguru_logger
    _HAVE_LOGURU = True
except Exception:  # pragma: no cover
    import logging

    logging.basicConfig(level=logging.INFO, format="%(asctime)s | %(levelname)s | %(message)s")
    logger = logging.getLogger("relora_b200")
    _HAVE_LOGURU = False


def silence_non_zero_rank(rank: int) -> None:
    """Only rank 0 logs (reference ``torchrun_main.py:371``)."""
    if rank == 0:
        return
    if _HAVE_LOGURU:
        logger.remove()
    else:  # pragma: no cover
        logger.setLevel(100)


class MetricSink:
    """Minimal run object: ``log``, ``config_update``, ``alert``, ``finish``; has ``id``/``name``."""

    def __init__(self, run_id: Optional[str] = None, name: Optional[str] = None):
        self.id = run_id or uuid.uuid4().hex[:8]
        self.name = name or f"run-{self.id}"

    def log(self, metrics: Dict[str, Any], step: Optional[int] = None) -> None:  # pragma: no cover
        pass

    def config_update(self, cfg: Dict[str, Any]) -> None:  # pragma: no cover
        pass

    def alert(self, title: str, text: str) -> None:
        logger.warning(f"[alert] {title}: {text}")

    def save(self, path: str) -> None:
        pass

    def watch(self, model, log_freq: int = 500) -> None:
        pass

    def finish(self) -> None:
        pass


def _to_jsonable(v):
    try:
        import torch

        if isinstance(v, torch.Tensor):
            return v.item() if v.numel() == 1 else v.detach().float().cpu().tolist()
    except Exception:  # pragma: no cover
        pass
    if isinstance(v, (set, tuple)):
        return list(v)
    return v


class JsonlSink(MetricSink):
    """Appends one JSON object per ``log`` call to ``<dir>/metrics.jsonl``."""

    def __init__(self, directory: Optional[str] = None, run_id: Optional[str] = None, name: Optional[str] = None):
        super().__init__(run_id, name)
        self._path = None
        self._pending = []
        if directory is not None:
            self.attach(directory)

    def attach(self, directory: str) -> None:
        os.makedirs(directory, exist_ok=True)
        self._path = os.path.join(directory, "metrics.jsonl")
        for rec in self._pending:
            self._write(rec)
        self._pending.clear()

    def _write(self, rec):
        with open(self._path, "a") as f:
            f.write(json.dumps(rec, default=str) + "\n")

    def log(self, metrics, step=None):
        rec = {k: _to_jsonable(v) for k, v in metrics.items()}
        rec["_step"] = step
        rec["_time"] = time.time()
        if self._path is None:
            self._pending.append(rec)
        else:
            self._write(rec)

    def config_update(self, cfg):
        self.log({"_config": {k: _to_jsonable(v) for k, v in cfg.items()}})


class WandbSink(MetricSink):
    def __init__(self, *, project: str, tags=None, run_id=None, notes=None):
        import wandb

        self._wandb = wandb
        wandb.init(project=project, tags=tags, id=run_id, resume="allow", notes=notes)
        super().__init__(wandb.run.id, wandb.run.name)

    def log(self, metrics, step=None):
        self._wandb.log(metrics, step=step)

    def config_update(self, cfg):
        self._wandb.config.update(cfg, allow_val_change=True)

    def alert(self, title, text):
        try:
            self._wandb.alert(title=title, text=text, level=self._wandb.AlertLevel.WARN)
        except Exception:  # pragma: no cover
            super().alert(title, text)

    def save(self, path):
        try:
            self._wandb.save(path, policy="now")
        except Exception:  # pragma: no cover
            pass

    def watch(self, model, log_freq=500):
        self._wandb.watch(model, log_freq=log_freq)

    def finish(self):
        self._wandb.finish()


def make_sink(*, project: str = "peft_pretraining", tags=None, run_id=None, notes=None, directory=None) -> MetricSink:
    """wandb when usable, JSONL otherwise.  ``RELORA_B200_NO_WANDB=1`` forces JSONL."""
    mode = os.environ.get("WANDB_MODE", "")
    if os.environ.get("RELORA_B200_NO_WANDB", "0") != "1" and mode != "disabled":
        try:
            return WandbSink(project=project, tags=tags, run_id=run_id, notes=notes)
        except Exception as e:  # no network, no API key, ...
            logger.warning(f"wandb unavailable ({type(e).__name__}: {e}); logging metrics to JSONL")
    return JsonlSink(directory, run_id=run_id)


def maybe_make_profiler(enabled: bool, run_name: str, rank: int):
    """torch.profiler with the reference's schedule (wait=1, warmup=1, active=3, repeat=2)."""
    if not enabled:
        return None
    import torch

    out = os.path.join("profiler_logs", str(run_name))
    prof = torch.profiler.profile(
        schedule=torch.profiler.schedule(wait=1, warmup=1, active=3, repeat=2),
        on_trace_ready=torch.profiler.tensorboard_trace_handler(out, worker_name=f"rank{rank}"),
        record_shapes=True,
        profile_memory=True,
        with_stack=True,
    )
    print(f"Rank {rank} profiling results will be saved to {out}")
    prof.start()
    return prof


@contextlib.contextmanager
def nvtx_range(name: str):
    """NVTX range when CUDA is present; a no-op otherwise (the reference has none)."""
    try:
        import torch

        on = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        on = False
    if on:
        torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        if on:
            torch.cuda.nvtx.range_pop()


class PhaseTimer:
    """Accumulates host wall time per named phase (save / eval / merge), like the reference's
    ad-hoc ``time.time()`` brackets, but queryable."""

    def __init__(self):
        self.totals: Dict[str, float] = {}

    @contextlib.contextmanager
    def phase(self, name: str):
        t0 = time.time()
        try:
            yield
        finally:
            self.totals[name] = self.totals.get(name, 0.0) + time.time() - t0
