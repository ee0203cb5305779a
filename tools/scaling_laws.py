"""Power-law fit of final loss against model size (reference ``notebooks/03_scaling_laws_plotting``, ``07_plotting``):
``L(N) = a · N^(-b) + c`` fitted by a grid over ``c`` and least squares in log space.

    python -m tools.scaling_laws --points 60e6:3.68 130e6:3.25 250e6:2.98 350e6:2.87 --predict 1.3e9
"""
from __future__ import annotations

import argparse
import math
from typing import List, Sequence, Tuple


def fit_power_law(points: Sequence[Tuple[float, float]]):
    """Returns ``(a, b, c, rmse)``."""
    xs = [math.log(n) for n, _ in points]
    best = None
    lo = min(l for _, l in points)
    for i in range(0, 400):
        c = lo * i / 400.0
        ys = [math.log(l - c) for _, l in points]
        n = len(xs)
        mx, my = sum(xs) / n, sum(ys) / n
        sxx = sum((x - mx) ** 2 for x in xs)
        if sxx == 0:
            continue
        slope = sum((x - mx) * (y - my) for x, y in zip(xs, ys)) / sxx
        icpt = my - slope * mx
        a, b = math.exp(icpt), -slope
        rmse = math.sqrt(sum((a * nn ** (-b) + c - l) ** 2 for nn, l in points) / n)
        if best is None or rmse < best[3]:
            best = (a, b, c, rmse)
    return best


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("--points", nargs="+", required=True, help="params:loss pairs")
    ap.add_argument("--predict", type=float, nargs="*", default=[])
    a = ap.parse_args(argv)
    pts: List[Tuple[float, float]] = [tuple(float(v) for v in p.split(":")) for p in a.points]
    aa, b, c, rmse = fit_power_law(pts)
    print(f"L(N) = {aa:.4g} * N^(-{b:.4f}) + {c:.4f}    rmse {rmse:.4g}")
    for n in a.predict:
        print(f"N = {n:.3g}: predicted loss {aa * n ** (-b) + c:.4f}")
    return aa, b, c, rmse


if __name__ == "__main__":
    main()
