"""Summarise / export the metrics of a run (the ``metrics.jsonl`` written next to the checkpoints when wandb is off; the role of
the reference's ``notebooks/07_plotting.ipynb`` loss and throughput plots).

    python -m tools.metrics checkpoints/run [--key loss] [--every 100] [--csv loss.csv] [--png loss.png]
"""
from __future__ import annotations

import argparse
import json
import os
from typing import Dict, List, Optional, Tuple


def load(path: str) -> List[dict]:
    f = os.path.join(path, "metrics.jsonl") if os.path.isdir(path) else path
    rows = []
    with open(f) as fh:
        for line in fh:
            line = line.strip()
            if line:
                rows.append(json.loads(line))
    return rows


def series(rows: List[dict], key: str) -> List[Tuple[int, float]]:
    """``(step, value)`` pairs of a scalar metric, in logging order."""
    out = []
    for r in rows:
        v = r.get(key)
        if isinstance(v, (int, float)) and r.get("_step") is not None:
            out.append((int(r["_step"]), float(v)))
    return out


def summarise(rows: List[dict]) -> Dict[str, Optional[float]]:
    loss = series(rows, "loss")
    tput = [v for _, v in series(rows, "throughput_tokens")]
    evals = series(rows, "final_eval_loss") or series(rows, "eval_loss")
    resets = [s for s, _ in series(rows, "n_lora_restarts")]
    tail = [v for _, v in loss[-max(1, len(loss) // 10):]]
    return {
        "logged_steps": float(len(loss)),
        "first_loss": loss[0][1] if loss else None,
        "last_loss": loss[-1][1] if loss else None,
        "mean_loss_last_10pct": sum(tail) / len(tail) if tail else None,
        "median_tokens_per_s": sorted(tput)[len(tput) // 2] if tput else None,
        "last_eval_loss": evals[-1][1] if evals else None,
        "lora_restarts": max((v for _, v in series(rows, "n_lora_restarts")), default=None) if resets else None,
    }


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("path")
    ap.add_argument("--key", default="loss")
    ap.add_argument("--every", type=int, default=1, help="keep every n-th point in the export")
    ap.add_argument("--csv", default=None)
    ap.add_argument("--png", default=None)
    a = ap.parse_args(argv)
    rows = load(a.path)
    print(json.dumps(summarise(rows), indent=1))
    pts = series(rows, a.key)[:: max(1, a.every)]
    if a.csv:
        with open(a.csv, "w") as fh:
            fh.write(f"step,{a.key}\n")
            fh.writelines(f"{s},{v:.8g}\n" for s, v in pts)
    if a.png:
        import matplotlib

        matplotlib.use("Agg")
        import matplotlib.pyplot as plt

        plt.figure(figsize=(8, 3))
        plt.plot([s for s, _ in pts], [v for _, v in pts])
        plt.xlabel("step")
        plt.ylabel(a.key)
        plt.tight_layout()
        plt.savefig(a.png, dpi=120)
    return pts


if __name__ == "__main__":
    main()
