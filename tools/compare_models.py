"""Logit comparison between two models / checkpoints on the same random token batch (reference ``notebooks/11_test_pythia``:
local SDPA Pythia vs HF, ``allclose(atol=1e-5)``; ``16_quantized``: L2 distance of the logits of a quantised model to the
full-precision one; ``12_test_relora_init``: wrapped vs unwrapped model at initialisation).

    python -m tools.compare_models --a checkpoints/run/model_5000 --b checkpoints/run/model_5000 --quantize_b 4bit
"""
from __future__ import annotations

import argparse
import json
import os

import torch


def load_model(path: str, quantize: str | None = None, device: str = "cpu", dtype=torch.float32):
    """A ReLoRA checkpoint directory (``relora_config.json``), a plain checkpoint directory or a config JSON (random init)."""
    from relora_b200.models import build_causal_lm, load_config
    from relora_b200.relora import ReLoRaModel

    if os.path.isdir(path) and os.path.exists(os.path.join(path, "relora_config.json")):
        m = ReLoRaModel.from_pretrained(path)
    else:
        cfg = load_config(os.path.join(path, "config.json") if os.path.isdir(path) else path)
        m = build_causal_lm(cfg)
        w = os.path.join(path, "pytorch_model.bin") if os.path.isdir(path) else None
        if w and os.path.exists(w):
            m.load_state_dict(torch.load(w, map_location="cpu", weights_only=True), strict=True)
    if quantize:
        m = ReLoRaModel(m if not isinstance(m, ReLoRaModel) else m.wrapped_model, r=8, lora_alpha=8, lora_dropout=0.0,
                        target_modules=["attn", "attention", "mlp"], quantize=quantize, keep_original_weights=True)
    return m.to(device=device, dtype=dtype).eval()


@torch.no_grad()
def compare_logits(model_a, model_b, vocab_size: int, batch: int = 2, seq: int = 64, seed: int = 0, device: str = "cpu") -> dict:
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, vocab_size, (batch, seq), generator=g).to(device)
    la = model_a(input_ids=ids, return_logits=True).logits.float()
    lb = model_b(input_ids=ids, return_logits=True).logits.float()
    d = la - lb
    return {"l2": float(d.norm()), "max_abs": float(d.abs().max()), "rel": float(d.norm() / la.norm().clamp(min=1e-30)),
            "allclose_1e-5": bool(torch.allclose(la, lb, atol=1e-5))}


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("--a", required=True)
    ap.add_argument("--b", required=True)
    ap.add_argument("--quantize_b", default=None, choices=[None, "4bit", "8bit", "nvfp4", "mxfp8"])
    ap.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--seq", type=int, default=64)
    a = ap.parse_args(argv)
    ma = load_model(a.a, None, a.device)
    mb = load_model(a.b, a.quantize_b, a.device)
    cfg = getattr(ma, "config", None) or ma.wrapped_model.config
    res = compare_logits(ma, mb, cfg.vocab_size, a.batch, a.seq, device=a.device)
    print(json.dumps(res, indent=1))
    return res


if __name__ == "__main__":
    main()
