"""Learning-rate schedule dump / plot (reference ``notebooks/04_plot_lr.ipynb``).

    python -m tools.plot_lr --scheduler cosine_restarts --num_training_steps 20000 --warmup_steps 500 \
        --cycle_length 5000 --restart_warmup_steps 100 --lr 1e-3 --csv lr.csv [--png lr.png]
"""
from __future__ import annotations

import argparse
from typing import List

import torch


def schedule(scheduler: str, num_training_steps: int, warmup_steps: int, lr: float = 1.0, min_lr_ratio: float = 0.1,
             cycle_length: int | None = None, restart_warmup_steps: int | None = None, adjust_step: int = 0) -> List[float]:
    """Learning rate at every update step, produced by the trainer's own scheduler factory."""
    from relora_b200.relora.schedulers import get_scheduler

    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=lr)
    sch = get_scheduler(opt, scheduler_type=scheduler, num_training_steps=num_training_steps, warmup_steps=warmup_steps,
                        min_lr_ratio=min_lr_ratio, cycle_length=cycle_length, restart_warmup_steps=restart_warmup_steps,
                        adjust_step=adjust_step)
    out = []
    for _ in range(num_training_steps):
        out.append(opt.param_groups[0]["lr"])
        opt.step()
        sch.step()
    return out


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("--scheduler", default="cosine_restarts", choices=["linear", "cosine", "cosine_restarts"])
    ap.add_argument("--num_training_steps", type=int, default=20000)
    ap.add_argument("--warmup_steps", type=int, default=500)
    ap.add_argument("--cycle_length", type=int, default=None)
    ap.add_argument("--restart_warmup_steps", type=int, default=None)
    ap.add_argument("--adjust_step", type=int, default=0)
    ap.add_argument("--min_lr_ratio", type=float, default=0.1)
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--csv", default=None)
    ap.add_argument("--png", default=None)
    a = ap.parse_args(argv)
    lrs = schedule(a.scheduler, a.num_training_steps, a.warmup_steps, a.lr, a.min_lr_ratio, a.cycle_length, a.restart_warmup_steps,
                   a.adjust_step)
    if a.csv:
        with open(a.csv, "w") as f:
            f.write("step,lr\n")
            f.writelines(f"{i},{v:.10g}\n" for i, v in enumerate(lrs))
    if a.png:
        import matplotlib  # optional dependency

        matplotlib.use("Agg")
        import matplotlib.pyplot as plt

        plt.figure(figsize=(8, 3))
        plt.plot(lrs)
        plt.xlabel("update step")
        plt.ylabel("learning rate")
        plt.tight_layout()
        plt.savefig(a.png, dpi=120)
    if not a.csv and not a.png:
        step = max(1, len(lrs) // 20)
        for i in range(0, len(lrs), step):
            print(f"{i:8d} {lrs[i]:.6g}")
    return lrs


if __name__ == "__main__":
    main()
