"""Sanity checks of a pre-tokenised dataset directory (reference ``notebooks/10_chunking``, ``14_check_pretokenization``,
``15_debug_dataloading``): ``args.json`` present and consistent, every sequence has exactly ``sequence_length`` ids inside the
vocabulary, optional decode of the first examples.

    python -m tools.check_dataset preprocessed_data/<dir> [--vocab_size 32100] [--samples 1000] [--tokenizer t5-base --show 2]
"""
from __future__ import annotations

import argparse
import json
import os


def check(path: str, vocab_size: int | None = None, samples: int = 1000) -> dict:
    import datasets

    with open(os.path.join(path, "args.json")) as f:
        prep = json.load(f)
    dd = datasets.load_from_disk(path)
    dd.set_format(type=None, columns=["input_ids"])
    seq = int(prep["sequence_length"])
    report = {"sequence_length": seq, "tokenizer": prep.get("tokenizer"), "splits": {}}
    for name in dd.keys() if hasattr(dd, "keys") else ["train"]:
        ds = dd[name] if hasattr(dd, "keys") else dd
        n = len(ds)
        k = min(n, samples)
        bad_len, lo, hi = 0, None, None
        for i in range(k):
            ids = ds[i * max(1, n // k) if k else 0]["input_ids"]
            bad_len += int(len(ids) != seq)
            mn, mx = min(ids), max(ids)
            lo = mn if lo is None else min(lo, mn)
            hi = mx if hi is None else max(hi, mx)
        ok = bad_len == 0 and (lo is None or lo >= 0) and (vocab_size is None or hi is None or hi < vocab_size)
        report["splits"][name] = {"sequences": n, "checked": k, "wrong_length": bad_len, "min_id": lo, "max_id": hi, "ok": ok,
                                  "tokens": n * seq}
    report["ok"] = all(s["ok"] for s in report["splits"].values())
    return report


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("path")
    ap.add_argument("--vocab_size", type=int, default=None)
    ap.add_argument("--samples", type=int, default=1000)
    ap.add_argument("--tokenizer", default=None)
    ap.add_argument("--show", type=int, default=0)
    a = ap.parse_args(argv)
    rep = check(a.path, a.vocab_size, a.samples)
    print(json.dumps(rep, indent=1))
    if a.show and a.tokenizer:
        import datasets
        from transformers import AutoTokenizer

        tok = AutoTokenizer.from_pretrained(a.tokenizer)
        ds = datasets.load_from_disk(a.path)["train"]
        for i in range(a.show):
            print(f"--- train[{i}] ---\n{tok.decode(ds[i]['input_ids'])[:500]}")
    return rep


if __name__ == "__main__":
    main()
