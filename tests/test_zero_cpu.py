"""ZeRO-1 (`--optimizer adam_zero`) and plain DP on CPU/gloo with 2 ranks: same parameters as a single rank with the
same global batch, consolidated optimizer state in the checkpoint."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = os.path.join(ROOT, "configs", "llama_9m.json")

WORKER = r'''
import os, sys, json, torch
sys.path.insert(0, sys.argv[1])
from torchrun_main import main
res = main(["--model_config", sys.argv[2], "--synthetic_data", "512", "--batch_size", "2", "--total_batch_size", "4", "--max_length", "16",
            "--lr", "1e-3", "--scheduler", "cosine", "--warmup_steps", "1", "--num_training_steps", "4", "--save_every", "4",
            "--eval_every", "100", "--save_dir", sys.argv[3], "--device", "cpu", "--dtype", "float32", "--workers", "0",
            "--use_peft", "--lora_r", "4", "--relora", "4", "--lora_dropout", "0.0", "--init_lora_a", "kaiming", "--optimizer", sys.argv[4]])
'''


def _run(tmp_path, nproc, opt, tag, port):
    d = str(tmp_path / tag)
    script = tmp_path / f"w_{tag}.py"
    script.write_text(WORKER)
    env = dict(os.environ, PYTHONPATH=ROOT, RELORA_B200_NO_WANDB="1", OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script), ROOT, CFG, d, opt]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    return d


@pytest.mark.slow
def test_zero1_matches_replicated_adam(tmp_path):
    a = _run(tmp_path, 2, "adam", "dp", 29711)
    b = _run(tmp_path, 2, "adam_zero", "zero", 29713)
    sa = torch.load(os.path.join(a, "model_4", "pytorch_model.bin"), weights_only=True)
    sb = torch.load(os.path.join(b, "model_4", "pytorch_model.bin"), weights_only=True)
    for k in sa:
        assert torch.allclose(sa[k], sb[k], atol=5e-6), k  # fp32 reduction-order noise (reduce_scatter vs all_reduce)
    oa = torch.load(os.path.join(a, "model_4", "optimizer.pt"), weights_only=False)["optimizer"]
    ob = torch.load(os.path.join(b, "model_4", "optimizer.pt"), weights_only=False)["optimizer"]
    assert set(oa["state"]) == set(ob["state"])  # consolidated: every parameter has state on rank 0
    for i in oa["state"]:
        assert torch.allclose(oa["state"][i]["exp_avg"], ob["state"][i]["exp_avg"], atol=1e-7)
