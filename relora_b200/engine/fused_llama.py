"""Whole-layer fused Llama + ReLoRA executor for B200 (stacked QKV / gate-up weights, LoRA folded into
the tcgen05 GEMMs, flat fp32 gradient accumulation, CUDA-graph captured micro-steps)."""
from __future__ import annotations


def supports(model, args):
    return False, "fused executor not built yet"


class FusedLlamaStepper:  # pragma: no cover - placeholder until the executor lands
    def __init__(self, *a, **k):
        raise NotImplementedError
