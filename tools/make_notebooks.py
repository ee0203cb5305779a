"""Generate ``notebooks/*.ipynb`` — the analysis notebooks of this repo (reference ``notebooks/01..16_*.ipynb``).

    python -m tools.make_notebooks [--out notebooks]

The upstream notebooks are scratch pads around a wandb account, the HF hub and local checkpoints.  These are written against this
repo's own modules / ``tools`` CLIs and run on CPU with synthetic inputs, so every one of them executes offline
(``tests/test_tools.py::test_notebooks_execute`` runs their code cells).  Point the path variables at real checkpoints / datasets
to reproduce the upstream analyses.
"""
from __future__ import annotations

import argparse
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SETUP = """import os, sys, json, math, tempfile
ROOT = os.path.abspath(os.path.join(os.getcwd(), "..")) if os.path.basename(os.getcwd()) == "notebooks" else os.getcwd()
sys.path.insert(0, ROOT)
os.environ.setdefault("RELORA_B200_NO_WANDB", "1")
import torch
torch.manual_seed(0)"""

TINY = """from relora_b200.models import LlamaForCausalLM, load_config
from relora_b200.relora import ReLoRaModel
cfg = load_config(os.path.join(ROOT, "configs", "llama_9m.json"))
def tiny_relora(r=8, **kw):
    torch.manual_seed(0)
    return ReLoRaModel(LlamaForCausalLM(cfg), r=r, lora_alpha=32, lora_dropout=0.0, target_modules=["attn", "attention", "mlp"],
                       keep_original_weights=True, init_lora_a="kaiming", **kw)"""

NOTEBOOKS = [
    ("01_peft_pretraining", "LoRA-only vs ReLoRA wrapping of a small Llama (upstream: first PEFT experiments with the `peft` package).",
     [TINY,
      """full = LlamaForCausalLM(cfg)
n_full = sum(p.numel() for p in full.parameters())
m = tiny_relora(r=8)
n_train = sum(p.numel() for p in m.parameters() if p.requires_grad)
print(f"full-rank parameters {n_full/1e6:.2f}M, trainable under ReLoRA r=8: {n_train/1e6:.2f}M ({100*n_train/n_full:.1f} %)")""",
      """ids = torch.randint(0, cfg.vocab_size - 1, (2, 32))
opt = torch.optim.AdamW([p for p in m.parameters() if p.requires_grad], lr=1e-3)
for step in range(3):
    loss = m(input_ids=ids, labels=ids).loss; loss.backward(); opt.step(); opt.zero_grad()
    print(step, float(loss))
m.merge_and_reinit(); print("merged; |B| =", float(sum(x.lora_B.weight.abs().sum() for x in m.modules() if hasattr(x, "lora_B"))))"""]),
    ("02_quick_debugs", "Parameter counts of the shipped model configs (upstream cell 0: llama_100m = 100.11712 M).",
     ["""from relora_b200.models import LlamaForCausalLM, load_config
for name in ("llama_9m", "llama_35m", "llama_100m"):
    c = load_config(os.path.join(ROOT, "configs", f"{name}.json"))
    with torch.device("meta"):
        n = sum(p.numel() for p in LlamaForCausalLM(c).parameters())
    print(f"{name}: {n/1e6:.5f} M parameters")"""]),
    ("03_scaling_laws_plotting", "Power-law fit of loss against model size (`tools.scaling_laws`).",
     ["""from tools.scaling_laws import fit_power_law
points = [(60e6, 3.68), (130e6, 3.25), (250e6, 2.98), (350e6, 2.87)]
fit = fit_power_law(points); print(fit)"""]),
    ("04_plot_lr", "The jagged cosine-with-restarts schedule of ReLoRA (`tools.plot_lr`, `relora_b200.relora.schedulers`).",
     ["""from tools.plot_lr import schedule
lrs = schedule("cosine_restarts", num_training_steps=100, warmup_steps=10, lr=1e-3, min_lr_ratio=0.1, cycle_length=25, restart_warmup_steps=5)
print([round(x * 1e3, 4) for x in lrs[:32]])
assert abs(lrs[25]) < 1e-12  # the update right before a merge runs at lr 0""",
      """try:
    import matplotlib.pyplot as plt
    plt.plot(lrs); plt.xlabel("update step"); plt.ylabel("lr"); plt.title("cosine_restarts")
except ImportError:
    print("matplotlib not installed: values printed above")"""]),
    ("05_check_ranks", "Singular values of the update learned by ReLoRA between two checkpoints (`tools.rank_analysis`).",
     [TINY,
      """from tools.rank_analysis import analyse
m = tiny_relora(r=4)
before = {k: v.clone() for k, v in m.wrapped_model.state_dict().items()}
ids = torch.randint(0, cfg.vocab_size - 1, (2, 32))
opt = torch.optim.AdamW([p for p in m.parameters() if p.requires_grad], lr=3e-3)
for cycle in range(3):              # three low-rank cycles -> rank of the accumulated update up to 12
    for _ in range(3):
        m(input_ids=ids, labels=ids).loss.backward(); opt.step(); opt.zero_grad()
    m.merge_and_reinit()
after = m.wrapped_model.state_dict()
for proj, stats in analyse(before, after, threshold=1e-4).items(): print(proj, stats)"""]),
    ("06_svd", "SVD of a single merged update: rank grows by r per ReLoRA cycle.",
     ["""r, out_f, in_f = 4, 48, 64
W = torch.zeros(out_f, in_f)
for cycle in range(3):
    B, A = torch.randn(out_f, r) * 0.1, torch.randn(r, in_f) * 0.1
    W += B @ A
    s = torch.linalg.svdvals(W)
    print(f"cycle {cycle + 1}: numerical rank {(s > 1e-5).sum().item()} (expected {r * (cycle + 1)})")"""]),
    ("07_plotting", "Loss / throughput curves of a run from its `metrics.jsonl` (`tools.metrics`; upstream pulls the same series from wandb).",
     ["""from tools.metrics import load, series, summarise
d = tempfile.mkdtemp()
with open(os.path.join(d, "metrics.jsonl"), "w") as f:
    for s in range(1, 41):
        f.write(json.dumps({"_step": s, "loss": 10.4 - 0.05 * s + 0.3 * (s % 10 == 1), "lr": 1e-3, "throughput_tokens": 4.1e5, "update_step": s}) + "\\n")
rows = load(d); print(summarise(rows)); print(series(rows, "loss")[:5])"""]),
    ("08_ranks_before_and_after", "Rank of the frozen weight's change before / after ReLoRA training vs plain LoRA (one cycle).",
     [TINY,
      """def delta_rank(n_cycles):
    m = tiny_relora(r=4)
    w0 = m.wrapped_model.model.layers[0].self_attn.q_proj.weight.detach().clone()
    ids = torch.randint(0, cfg.vocab_size - 1, (2, 32)); opt = torch.optim.AdamW([p for p in m.parameters() if p.requires_grad], lr=3e-3)
    for _ in range(n_cycles):
        for _ in range(2): m(input_ids=ids, labels=ids).loss.backward(); opt.step(); opt.zero_grad()
        m.merge_and_reinit()
    dw = m.wrapped_model.model.layers[0].self_attn.q_proj.weight.detach() - w0
    return int((torch.linalg.svdvals(dw.float()) > 1e-6).sum())
print("LoRA (1 cycle):", delta_rank(1), " ReLoRA (4 cycles):", delta_rank(4))"""]),
    ("09_bar_plot", "Count of small singular values per projection type (upstream: bar plot of #σ < 0.1 for Q/K/V/up/down).",
     ["""import collections
counts = collections.Counter()
for name, shape in {"q_proj": (64, 64), "k_proj": (64, 64), "v_proj": (64, 64), "up_proj": (172, 64), "down_proj": (64, 172)}.items():
    dW = torch.randn(shape[0], 8) @ torch.randn(8, shape[1]) * 0.05      # a rank-8 update
    counts[name] = int((torch.linalg.svdvals(dW) < 0.1).sum())
print(dict(counts))"""]),
    ("10_chunking", "Concatenate-and-chunk tokenisation (`relora_b200.data.hf_disk.tokenize_and_chunk` semantics: remainder dropped).",
     ["""docs = [[5, 6, 7, 1], [8, 9, 1], [10, 11, 12, 13, 14, 1]]           # token ids with an EOS (1) appended per document
flat = [t for d in docs for t in d]
L = 4
chunks = [flat[i:i + L] for i in range(0, len(flat) - len(flat) % L, L)]
print(chunks, "dropped tail:", flat[len(flat) - len(flat) % L:])"""]),
    ("11_test_pythia", "GPT-NeoX / Pythia implementation vs the Hugging Face one on random weights (upstream: allclose(atol=1e-5) on pythia-1b).",
     ["""from relora_b200.models import GPTNeoXForCausalLM, SimpleConfig
c = SimpleConfig(model_type="gpt_neox", vocab_size=512, hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=256,
                 rotary_pct=0.25, max_position_embeddings=128, layer_norm_eps=1e-5, use_parallel_residual=True, hidden_act="gelu",
                 rotary_emb_base=10000, tie_word_embeddings=False)
ours = GPTNeoXForCausalLM(c).eval()
ids = torch.randint(0, 512, (2, 17))
try:
    from transformers import GPTNeoXConfig, GPTNeoXForCausalLM as HF
    hf = HF(GPTNeoXConfig(**{k: v for k, v in c.to_dict().items() if k != "model_type"})).eval()
    ours.load_hf_state_dict(hf.state_dict(), strict=True)
    print("max |logit diff| vs HF:", float((ours(input_ids=ids).logits - hf(input_ids=ids).logits).abs().max()))
except Exception as e:
    print("transformers GPT-NeoX not usable here:", type(e).__name__, e); print(ours(input_ids=ids).logits.shape)"""]),
    ("12_test_relora_init", "Wrap a Pythia model in ReLoRA and check the loss (upstream: pythia-1.4b, loss 4.3360 on one sentence).",
     ["""from relora_b200.models import GPTNeoXForCausalLM, SimpleConfig
from relora_b200.relora import ReLoRaModel
c = SimpleConfig(model_type="gpt_neox", vocab_size=512, hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=256,
                 rotary_pct=0.25, max_position_embeddings=128, tie_word_embeddings=False)
base = GPTNeoXForCausalLM(c)
ids = torch.randint(0, 512, (1, 24))
l0 = float(base(input_ids=ids, labels=ids).loss)
m = ReLoRaModel(base, r=8, lora_alpha=32, lora_dropout=0.0, target_modules=["attn", "attention", "mlp"], keep_original_weights=True)
print("loss before / after wrapping (B = 0 -> identical):", l0, float(m(input_ids=ids, labels=ids).loss))
print([n for n, _ in m.named_parameters() if "lora_A" in n][:4])"""]),
    ("13_zero_optimizer_resets", "What an optimizer reset leaves in the Adam moments (`relora_b200.relora.optimizer_reset`, `relora_b200.utils`).",
     ["""from relora_b200.relora import optimizer_reset
from relora_b200.utils import optimizer_state_size
p = [torch.nn.Parameter(torch.randn(64, 32)) for _ in range(2)]
opt = torch.optim.AdamW(p, lr=1e-3)
for q in p: q.grad = torch.randn_like(q)
opt.step(); print("before:", optimizer_state_size(opt))
pct = optimizer_reset(opt, reset_params=p, optimizer_state_keys=["exp_avg", "exp_avg_sq"], reset_optimizer_on_relora=False,
                      optimizer_random_pruning=0.0, optimizer_magnitude_pruning=0.9)
print(f"zeroed {pct:.1f} %; after:", optimizer_state_size(opt))"""]),
    ("14_check_pretokenization", "Sanity check of a pre-tokenised dataset directory (`tools.check_dataset`).",
     ["""from relora_b200.data.synthetic import write_synthetic_hf_dataset
from tools.check_dataset import check
d = os.path.join(tempfile.mkdtemp(), "synthetic_t5-base_64")
write_synthetic_hf_dataset(d, n_train=64, n_val=8, seq_len=64, vocab_size=32100)
print(check(d, vocab_size=32100, samples=32))"""]),
    ("15_debug_dataloading", "Megatron / NeoX mmap dataset -> GPT2Dataset samples of seq_length + 1 tokens (upstream cell 6: batch [1024, 2049]).",
     ["""import numpy as np
from relora_b200.data.neox import GPT2Dataset, MMapIndexedDatasetBuilder, make_dataset
d = tempfile.mkdtemp(); prefix = os.path.join(d, "toy_text_document")
b = MMapIndexedDatasetBuilder(prefix + ".bin", dtype=np.uint16)
rng = np.random.RandomState(0)
for _ in range(50): b.add_item(torch.from_numpy(rng.randint(0, 1000, size=rng.randint(20, 80)).astype(np.int64))); b.end_document()
b.finalize(prefix + ".idx")
ds = make_dataset(prefix, "mmap")
g = GPT2Dataset("toy", prefix, np.arange(len(ds), dtype=np.int32), ds, num_samples=16, seq_length=32, seed=1234, build_index_mappings=True)
print(len(g), g[0]["input_ids"].shape)   # every sample carries seq_length + 1 tokens"""]),
    ("16_quantized", "Block-scaled (MXFP8 / NVFP4) frozen weights: logit distance to the full-precision model and resident bytes "
     "(upstream: bitsandbytes NF4, L2 distance 83.9 on llama_250m).",
     [TINY,
      """from relora_b200.utils import frozen_weight_bytes
ids = torch.randint(0, cfg.vocab_size - 1, (2, 32))
ref = tiny_relora(r=8)
for q in ("8bit", "4bit"):
    m = tiny_relora(r=8, quantize=q)
    m.wrapped_model.load_state_dict(ref.wrapped_model.state_dict())
    d = float((m(input_ids=ids).logits - ref(input_ids=ids).logits).norm())
    b = frozen_weight_bytes(m)
    print(f"{q}: L2 distance of logits {d:.3f}; frozen weights resident {b['resident_bytes']/b['bf16_bytes']:.3f} x bf16")"""]),
]


def notebook(title: str, blurb: str, cells):
    nb_cells = [{"cell_type": "markdown", "metadata": {}, "source": [f"# {title}\n", "\n", blurb + "\n"]},
                {"cell_type": "code", "metadata": {}, "execution_count": None, "outputs": [], "source": SETUP.splitlines(keepends=True)}]
    for c in cells:
        nb_cells.append({"cell_type": "code", "metadata": {}, "execution_count": None, "outputs": [], "source": c.splitlines(keepends=True)})
    return {"cells": nb_cells, "metadata": {"kernelspec": {"display_name": "Python 3", "language": "python", "name": "python3"},
                                            "language_info": {"name": "python"}}, "nbformat": 4, "nbformat_minor": 5}


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "notebooks"))
    a = ap.parse_args(argv)
    os.makedirs(a.out, exist_ok=True)
    for name, blurb, cells in NOTEBOOKS:
        with open(os.path.join(a.out, name + ".ipynb"), "w") as f:
            json.dump(notebook(name, blurb, cells), f, indent=1)
            f.write("\n")
    print(f"wrote {len(NOTEBOOKS)} notebooks to {a.out}")


if __name__ == "__main__":
    main()
