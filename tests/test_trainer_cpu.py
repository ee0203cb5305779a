"""End-to-end trainer tests on CPU / gloo (BASELINE.json config 1: llama_9m plumbing)."""
import json
import os
import subprocess
import sys

import pytest
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = os.path.join(ROOT, "configs", "llama_9m.json")


def _args(save_dir, extra=()):
    return [
        "--model_config", CFG, "--synthetic_data", "2048", "--batch_size", "2", "--total_batch_size", "4",
        "--max_length", "32", "--lr", "1e-3", "--use_peft", "--lora_r", "4", "--relora", "4", "--cycle_length", "4",
        "--restart_warmup_steps", "1", "--scheduler", "cosine_restarts", "--warmup_steps", "2",
        "--num_training_steps", "8", "--save_every", "4", "--eval_every", "100", "--save_dir", save_dir,
        "--device", "cpu", "--dtype", "float32", "--workers", "0", *extra,
    ]


def test_full_rank_warmup_cpu(tmp_path):
    """llama_9m full-rank warm-up on CPU/gloo world_size=1, synthetic tokens."""
    from torchrun_main import main

    d = str(tmp_path / "warm")
    res = main(["--model_config", CFG, "--synthetic_data", "1024", "--batch_size", "2", "--total_batch_size", "2",
                "--max_length", "32", "--lr", "2e-3", "--scheduler", "cosine", "--warmup_steps", "2",
                "--num_training_steps", "6", "--save_every", "3", "--eval_every", "100", "--save_dir", d,
                "--device", "cpu", "--dtype", "float32", "--workers", "0"])
    assert res["update_step"] == 6 and res["executor"] == "ModuleStepper"
    assert sorted(x for x in os.listdir(d) if x.startswith("model_")) == ["model_3", "model_6"]
    assert os.path.exists(os.path.join(d, "training_config.yaml"))
    assert res["final_eval_loss"] < 10.6


def test_relora_run_checkpoint_layout_and_autoresume(tmp_path):
    from torchrun_main import main

    d = str(tmp_path / "relora")
    res = main(_args(d))
    assert res["update_step"] == 8 and res["n_lora_restarts"] == 1 and res["n_optimizer_resets"] == 1
    m4 = os.path.join(d, "model_4")
    assert sorted(os.listdir(m4)) == ["config.json", "optimizer.pt", "pytorch_model.bin", "relora_config.json",
                                      "training_state.json"]
    ts = json.load(open(os.path.join(m4, "training_state.json")))
    assert set(ts) == {"global_step", "update_step", "tokens_seen", "tokens_seen_before", "n_lora_restarts",
                       "n_optimizer_resets", "update_time", "wandb_id"}
    assert ts["update_step"] == 4 and ts["global_step"] == 8 and ts["tokens_seen"] == 4 * 4 * 32
    oc = torch.load(os.path.join(m4, "optimizer.pt"), weights_only=False)
    assert set(oc) == {"optimizer", "scheduler", "update_step", "global_step", "config", "dtype"}
    assert set(oc["optimizer"]) == {"state", "param_groups"}
    st0 = oc["optimizer"]["state"][0]
    assert set(st0) == {"step", "exp_avg", "exp_avg_sq"}
    # an existing save_dir without --autoresume is an error
    with pytest.raises(ValueError):
        main(_args(d))
    # autoresume continues from model_8 and trains to 12
    res2 = main(_args(d, ("--autoresume", "true", "--num_training_steps", "12")))
    assert res2["update_step"] == 12
    assert "model_12" in os.listdir(d)
    cfg = yaml.safe_load(open(os.path.join(d, "training_config.yaml")))
    assert cfg["num_training_steps"] == 12 and cfg["use_peft"] is True


def test_warm_start_then_relora_and_keep_checkpoints(tmp_path):
    from torchrun_main import main

    warm = str(tmp_path / "warm")
    main(["--model_config", CFG, "--synthetic_data", "1024", "--batch_size", "2", "--total_batch_size", "2",
          "--max_length", "32", "--lr", "1e-3", "--scheduler", "cosine", "--warmup_steps", "1",
          "--num_training_steps", "4", "--save_every", "4", "--eval_every", "100", "--save_dir", warm,
          "--device", "cpu", "--dtype", "float32", "--workers", "0"])
    d = str(tmp_path / "relora")
    res = main(_args(d, ("--warmed_up_model", os.path.join(warm, "model_4"), "--num_training_steps", "12",
                         "--optimizer_magnitude_pruning", "0.9", "--reset_optimizer_on_relora", "false",
                         "--keep_checkpoints", "1", "--save_every", "4")))
    # scheduler starts at the warm-start step (4): 8 scheduled steps, resets at relative step 5
    assert res["update_step"] == 12 and res["n_lora_restarts"] == 1
    assert [x for x in sorted(os.listdir(d)) if x.startswith("model_")] == ["model_12"]


def test_yaml_training_config(tmp_path):
    from relora_b200.config import parse_args

    y = tmp_path / "cfg.yaml"
    y.write_text(yaml.safe_dump({"model_config": CFG, "synthetic_data": "64", "batch_size": 2, "lr": "4e-4",
                                 "use_peft": True, "relora": 10, "scheduler": "cosine_restarts",
                                 "restart_warmup_steps": 2, "num_training_steps": 20, "max_train_tokens": None}))
    args = parse_args(["--training_config", str(y), "--device", "cpu"])
    assert args.lr == 4e-4 and args.cycle_length == 10 and args.use_peft and args.total_batch_size == 2
    with pytest.raises(RuntimeError):
        parse_args(["--training_config", str(y), "--lr", "1e-3"])


def test_cli_bool_flags_and_validation():
    from relora_b200.config import parse_args

    base = ["--model_config", CFG, "--synthetic_data", "8", "--batch_size", "2"]
    a = parse_args(base + ["--use_peft", "--relora", "5"])
    assert a.use_peft is True and a.cycle_length == 5
    a = parse_args(base + ["--use_peft", "false"])
    assert a.use_peft is False and a.relora is None and a.lora_r is None
    a = parse_args(base + ["--relora", "7"])  # relora implies use_peft
    assert a.use_peft is True
    a = parse_args(base + ["--max_train_tokens", "1M", "--total_batch_size", "4"])
    assert a.num_training_steps == 250_000  # tokens // sequences quirk
    a = parse_args(base + ["--skip_batches", "3,9"])
    assert a.skip_batches == {3, 9}
    with pytest.raises(ValueError):
        parse_args(["--model_config", CFG, "--batch_size", "2"])  # no data source
    with pytest.raises(ValueError):
        parse_args(base + ["--use_peft", "--optimizer_random_pruning", "0.5", "--optimizer_magnitude_pruning", "0.5",
                           "--reset_optimizer_on_relora", "false"])
    with pytest.raises(NotImplementedError):
        parse_args(base + ["--dtype", "fp16"])


def test_fsdp_is_refused(tmp_path):
    from torchrun_main import main

    with pytest.raises(RuntimeError, match="FSDP is not supported"):
        main(_args(str(tmp_path / "x"), ("--distributed_type", "fsdp")))


@pytest.mark.slow
def test_two_rank_gloo_matches_single_rank(tmp_path):
    """DDP semantics on CPU: 2 ranks x batch 2 == 1 rank x batch 2 x accumulation 2 (same global batch)."""
    env = dict(os.environ, PYTHONPATH=ROOT, RELORA_B200_NO_WANDB="1", OMP_NUM_THREADS="2")
    d2 = str(tmp_path / "ws2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "torchrun_main.py"), *_args(d2, ("--lora_dropout", "0.0"))]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    assert "model_8" in os.listdir(d2)
    ts = json.load(open(os.path.join(d2, "model_8", "training_state.json")))
    assert ts["update_step"] == 8 and ts["global_step"] == 8  # ga = 4 / (2*2) = 1


def test_flat_store_padded_storage_matches_torch_adamw():
    """A 2-D parameter stored in a larger zero-padded block (llama_1b's 5461-wide MLP on the fused path) trains exactly
    like the unpadded tensor; moments / checkpoints / pruning only ever see the logical tensor."""
    import torch
    from relora_b200.parallel.flat import FlatAdamW, FlatParamStore

    torch.manual_seed(0)
    a = torch.nn.Parameter(torch.randn(4, 5))
    b = torch.nn.Parameter(torch.randn(7, 3))
    c = torch.nn.Parameter(torch.randn(6))
    ra, rb_, rc = (torch.nn.Parameter(t.detach().clone()) for t in (a, b, c))
    store = FlatParamStore([("a", a), ("b", b), ("c", c)], world_size=1, grad_dtype=torch.float32,
                           storage_shapes={id(a): (4, 8), id(b): (8, 3)})
    assert store.is_padded(a) and not a.is_contiguous() and b.is_contiguous()
    opt = FlatAdamW(store, lr=1e-2, weight_decay=0.1)
    ref = torch.optim.AdamW([ra, rb_, rc], lr=1e-2, weight_decay=0.1)
    for it in range(3):
        for p, q in ((a, ra), (b, rb_), (c, rc)):
            gr = torch.randn_like(q)
            q.grad = gr.clone()
            p.grad.copy_(gr)
        opt.step()
        ref.step()
        opt.zero_grad()
    for p, q in ((a, ra), (b, rb_), (c, rc)):
        assert torch.allclose(p, q, atol=1e-6)
    o, n = store.segment(a)
    blk = store.params[o:o + n].view(4, 8)
    assert n == 32 and torch.count_nonzero(blk[:, 5:]) == 0  # the padding never moves
    sd = opt.state_dict()
    rsd = ref.state_dict()
    for i in range(3):
        assert torch.allclose(sd["state"][i]["exp_avg"], rsd["state"][i]["exp_avg"], atol=1e-6)
        assert sd["state"][i]["exp_avg"].shape == rsd["state"][i]["exp_avg"].shape
    # round trip through the torch layout
    opt2 = FlatAdamW(store, lr=1e-2, weight_decay=0.1)
    opt2.load_state_dict(sd)
    assert torch.equal(opt2.exp_avg, opt.exp_avg) and torch.equal(opt2.exp_avg_sq, opt.exp_avg_sq)
    # magnitude pruning takes its quantile over the logical tensor, not the padded block
    from relora_b200.relora.optim_reset import magnitude_pruning_

    want = opt.state[a]["exp_avg"].clone()
    magnitude_pruning_(want, 0.5)
    opt.prune_state([a], ["exp_avg"], "magnitude", 0.5)
    assert torch.equal(opt.state[a]["exp_avg"], want)
