"""Analysis tools (tools/): the role of the reference's notebooks, tested on CPU."""
import json
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lr_schedule_dump(tmp_path):
    from tools.plot_lr import main, schedule

    lrs = schedule("cosine_restarts", 200, 10, lr=1.0, cycle_length=50, restart_warmup_steps=5)
    assert len(lrs) == 200 and lrs[0] == 0.0 and max(lrs) <= 1.0 + 1e-9
    # jagged: the rate drops to ~0 right after every restart boundary and climbs again
    assert lrs[50] < 0.05 and lrs[56] > lrs[50]
    out = tmp_path / "lr.csv"
    main(["--scheduler", "cosine", "--num_training_steps", "20", "--warmup_steps", "2", "--csv", str(out)])
    assert out.read_text().splitlines()[0] == "step,lr" and len(out.read_text().splitlines()) == 21


def test_rank_analysis_counts_low_rank_update():
    from tools.rank_analysis import analyse, fold_lora

    torch.manual_seed(0)
    w0 = torch.randn(64, 48)
    b, a = torch.randn(64, 4), torch.randn(4, 48)
    sd_after = {"wrapped_model.model.layers.0.self_attn.q_proj.weight": w0.clone(),
                "wrapped_model.model.layers.0.self_attn.q_proj.lora_A.weight": a,
                "wrapped_model.model.layers.0.self_attn.q_proj.lora_B.weight": b,
                "wrapped_model.model.norm.weight": torch.ones(48)}
    after = fold_lora(sd_after, scale=0.5)
    before = {"model.layers.0.self_attn.q_proj.weight": w0}
    res = analyse(before, after, threshold=1e-3)
    q = res["q_proj"]
    assert q["singular_values"] == 48 and q["singular_values_below_threshold"] == 44  # rank-4 update
    assert 1.0 < q["mean_effective_rank"] <= 4.0 + 1e-3


def test_scaling_law_fit_recovers_parameters():
    from tools.scaling_laws import fit_power_law

    pts = [(n, 400.0 * n ** -0.3 + 1.8) for n in (6e7, 1.3e8, 2.5e8, 3.5e8, 1.3e9)]
    a, b, c, rmse = fit_power_law(pts)
    assert abs(b - 0.3) < 0.03 and abs(c - 1.8) < 0.1 and rmse < 1e-2


def test_compare_models_wrapped_vs_plain(tmp_path):
    from relora_b200.models import build_causal_lm, load_config
    from relora_b200.relora import ReLoRaModel
    from tools.compare_models import compare_logits

    cfg = load_config(os.path.join(ROOT, "configs", "llama_9m.json"))
    torch.manual_seed(0)
    m = build_causal_lm(cfg).eval()
    import copy

    w = ReLoRaModel(copy.deepcopy(m), r=8, lora_alpha=8, lora_dropout=0.0, target_modules=["attn", "mlp"]).eval()
    res = compare_logits(m, w, cfg.vocab_size, batch=2, seq=16)
    assert res["allclose_1e-5"] and res["l2"] < 1e-3  # B = 0 at initialisation: the wrapped model is the same function


def test_check_dataset_reports_sizes(tmp_path):
    import datasets

    from tools.check_dataset import check

    seq = 8
    d = datasets.DatasetDict({"train": datasets.Dataset.from_dict({"input_ids": [list(range(i, i + seq)) for i in range(20)]}),
                              "validation": datasets.Dataset.from_dict({"input_ids": [list(range(seq))]})})
    p = str(tmp_path / "ds")
    d.save_to_disk(p)
    json.dump({"sequence_length": seq, "tokenizer": "none"}, open(os.path.join(p, "args.json"), "w"))
    rep = check(p, vocab_size=100)
    assert rep["ok"] and rep["splits"]["train"]["sequences"] == 20 and rep["splits"]["train"]["tokens"] == 160
    assert not check(p, vocab_size=10)["ok"]


def test_metrics_summary_of_a_real_cpu_run(tmp_path):
    """tools.metrics reads the JSONL sink the trainer writes when wandb is off."""
    from tools.metrics import load, main, series, summarise
    from torchrun_main import main as train

    d = str(tmp_path / "run")
    train(["--model_config", os.path.join(ROOT, "configs", "llama_9m.json"), "--synthetic_data", "512", "--batch_size", "2",
           "--total_batch_size", "2", "--max_length", "16", "--lr", "1e-3", "--scheduler", "cosine", "--warmup_steps", "1",
           "--num_training_steps", "5", "--save_every", "5", "--eval_every", "100", "--save_dir", d, "--device", "cpu",
           "--dtype", "float32", "--workers", "0"])
    rows = load(d)
    losses = series(rows, "loss")
    assert len(losses) == 5 and all(v == v for _, v in losses)
    s = summarise(rows)
    assert s["logged_steps"] == 5 and s["first_loss"] > 5.0 and s["last_eval_loss"] is not None
    out = tmp_path / "loss.csv"
    main([d, "--csv", str(out)])
    assert out.read_text().splitlines()[0] == "step,loss" and len(out.read_text().splitlines()) == 6
