"""Multi-GPU tests (NCCL + NVLink peer memory).  Need >= 2 GPUs: `gpurun --gpus 2 -- pytest -m gpu tests/test_multigpu.py`."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(mode, nproc=2, port=29641, timeout=420):
    env = dict(os.environ, PYTHONPATH=ROOT, RELORA_B200_NO_WANDB="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py"), mode]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout)
    log_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(log_dir, exist_ok=True)
    with open(os.path.join(log_dir, f"multigpu_{mode}.log"), "w") as f:
        f.write(out.stdout + "\n=== stderr ===\n" + out.stderr)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-4000:])
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


def _nproc():
    return min(torch.cuda.device_count(), 8)


def test_peer_memory_allreduce_matches_nccl():
    res = _run("allreduce", nproc=_nproc(), port=29641)
    assert any(k.startswith("p2p_") for k in res)


def test_fused_update_p2p_matches_nccl_training():
    a = _run("train_nccl", nproc=_nproc(), port=29643)
    b = _run("train_p2p", nproc=_nproc(), port=29645)
    assert a["transport"] == "nccl" and b["transport"] == "p2p"
    assert a["executor"] == b["executor"] == "FusedLlamaStepper"
    for x, y in zip(a["losses"], b["losses"]):
        assert abs(x - y) < 5e-2, (a, b)
    # same data, same seeds: parameters agree up to bf16 reduction-order noise
    assert abs(a["checksum"][1] - b["checksum"][1]) / abs(a["checksum"][1]) < 1e-3, (a, b)


@pytest.mark.parametrize("transport", ["p2p", "nccl"])
def test_sixty_updates_stay_finite_and_replicated(transport):
    """60 updates at world = all visible GPUs through two ReLoRA resets: finite losses / norms / parameters, identical replicas."""
    res = _run(f"long_{transport}", nproc=_nproc(), port=29651 if transport == "p2p" else 29653, timeout=900)
    assert res["transport"] == transport and res["executor"] == "FusedLlamaStepper"
    assert res["restarts"] >= 2 and res["steps"] == 60
    assert all(l == l for l in res["losses"])


def test_module_path_uses_the_peer_memory_update():
    """Full-rank Llama (module path, no fused executor): the hand-written NVLink update is the transport, and it tracks NCCL."""
    a = _run("module_p2p", nproc=_nproc(), port=29655, timeout=900)
    b = _run("module_nccl", nproc=_nproc(), port=29657, timeout=900)
    assert a["transport"] == "p2p" and b["transport"] == "nccl"
    assert a["executor"] == b["executor"] == "ModuleStepper"
    assert abs(a["losses"][-1] - b["losses"][-1]) < 0.15, (a, b)
