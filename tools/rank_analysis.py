"""Singular-value / rank analysis of the update learned between two checkpoints (reference ``notebooks/05_check_ranks``,
``06_svd``, ``08_ranks_before_and_after``, ``09_bar_plot``: "singular values < 0.1 of the learned ΔW, summed over layers").

    python -m tools.rank_analysis --before checkpoints/warmup/model_5000 --after checkpoints/relora/model_20000 [--threshold 0.1]

Works on ``pytorch_model.bin`` files (or directories containing one) in the reference checkpoint layout; LoRA factors that are
still un-merged in ``--after`` are folded in (``W + s·B·A``) when ``relora_config.json`` is present.
"""
from __future__ import annotations

import argparse
import json
import os
import re
from collections import defaultdict
from typing import Dict

import torch

PROJ = ("q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj", "query_key_value", "dense", "dense_h_to_4h",
        "dense_4h_to_h")


def load_state(path: str) -> Dict[str, torch.Tensor]:
    f = os.path.join(path, "pytorch_model.bin") if os.path.isdir(path) else path
    sd = torch.load(f, map_location="cpu", weights_only=True)
    scale = None
    cfg = os.path.join(os.path.dirname(f), "relora_config.json")
    if os.path.exists(cfg):
        c = json.load(open(cfg))
        scale = c["lora_alpha"] / c["r"]
    return fold_lora(sd, scale)


def fold_lora(sd: Dict[str, torch.Tensor], scale: float | None) -> Dict[str, torch.Tensor]:
    """``{name.weight: W (+ s·B·A)}`` for every 2-D weight (LoRA factors removed)."""
    out = {}
    for k, v in sd.items():
        if ".lora_A." in k or ".lora_B." in k or v.dim() != 2:
            continue
        w = v.float()
        a, b = k.replace(".weight", ".lora_A.weight"), k.replace(".weight", ".lora_B.weight")
        if scale is not None and a in sd and b in sd:
            w = w + scale * (sd[b].float() @ sd[a].float())
        out[k.replace("wrapped_model.", "")] = w
    return out


def analyse(before: Dict[str, torch.Tensor], after: Dict[str, torch.Tensor], threshold: float = 0.1):
    """Per projection type: number of singular values of ΔW below ``threshold`` (summed over layers), total count and the
    mean effective rank ``exp(H(σ/Σσ))``."""
    stats = defaultdict(lambda: {"below": 0, "total": 0, "eff_rank": [], "layers": 0})
    for k, wa in after.items():
        if k not in before or before[k].shape != wa.shape:
            continue
        m = re.search("|".join(PROJ), k)
        if not m:
            continue
        s = torch.linalg.svdvals(wa - before[k])
        st = stats[m.group(0)]
        st["below"] += int((s < threshold).sum())
        st["total"] += s.numel()
        p = s / s.sum().clamp(min=1e-30)
        st["eff_rank"].append(float(torch.exp(-(p * torch.log(p.clamp(min=1e-30))).sum())))
        st["layers"] += 1
    return {k: {"singular_values_below_threshold": v["below"], "singular_values": v["total"], "layers": v["layers"],
                "mean_effective_rank": sum(v["eff_rank"]) / max(1, len(v["eff_rank"]))} for k, v in stats.items()}


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("--before", required=True)
    ap.add_argument("--after", required=True)
    ap.add_argument("--threshold", type=float, default=0.1)
    ap.add_argument("--json", default=None)
    a = ap.parse_args(argv)
    res = analyse(load_state(a.before), load_state(a.after), a.threshold)
    for k, v in sorted(res.items()):
        print(f"{k:18s} σ<{a.threshold}: {v['singular_values_below_threshold']:7d} / {v['singular_values']:7d}   "
              f"layers {v['layers']:3d}   effective rank {v['mean_effective_rank']:.1f}")
    if a.json:
        json.dump(res, open(a.json, "w"), indent=1)
    return res


if __name__ == "__main__":
    main()
