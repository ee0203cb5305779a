"""Multi-GPU worker (launched by tests/test_multigpu.py through torch.distributed.run)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("RELORA_B200_NO_WANDB", "1")
BF = torch.bfloat16


def check_allreduce(comm, rank, world):
    """two-shot peer-memory all-reduce (P2P and multicast variants) vs NCCL, sizes 64 KB .. 64 MB"""
    out = {}
    for use_mc in (False, True):
        comm.use_multicast = use_mc
        for n in (32 * 1024, 1024 * 1024, 8 * 1024 * 1024 + 8 * world, 32 * 1024 * 1024):
            n = (n // (8 * world)) * 8 * world
            buf = comm.alloc(n, BF)
            if use_mc and not buf.mc_base:
                out["multicast"] = "unavailable"
                continue
            g = torch.Generator(device="cuda").manual_seed(100 + rank)
            x = (torch.randn(n, device="cuda", generator=g) * 0.5).to(BF)
            buf.tensor.copy_(x)
            ref = x.float().clone()
            dist.all_reduce(ref)
            torch.cuda.synchronize()
            dist.barrier()
            comm.all_reduce_(buf)
            torch.cuda.synchronize()
            err = float((buf.tensor.float() - ref).abs().max() / ref.abs().max())
            assert err < 2e-2, (use_mc, n, err)
            out[f"{'mc' if use_mc else 'p2p'}_{n}"] = err
    comm.use_multicast = True
    return out


def check_training(rank, world, transport):
    """3 updates of the fused executor with DP=world; returns losses + a parameter checksum"""
    from relora_b200.engine.api import TrainingEngine
    from relora_b200.ops import fused
    from relora_b200.parallel.dist import init_distributed

    info = init_distributed("cuda", "nccl")
    eng = TrainingEngine.build(info, model_config=os.path.join(ROOT, "configs", "llama_35m.json"), batch_size=2,
                               gradient_accumulation=2, total_batch_size=4 * world, max_length=128, use_peft=True, lora_r=128,
                               relora=1000, cycle_length=1000, scheduler="cosine_restarts", warmup_steps=2, restart_warmup_steps=1,
                               lr=1e-3, num_training_steps=1000, dtype="bfloat16", device="cuda", init_lora_a="kaiming",
                               comm=transport, seed=0)
    fused.seed_state.set(info.device, 777)
    g = torch.Generator().manual_seed(5 + rank)
    losses = []
    for _ in range(3):
        ids = torch.randint(0, 32000, (2, 2, 128), generator=g).cuda()
        losses.append(float(eng.train_step_device(ids)))
    p = eng.stepper.store.params.float()
    chk = [float(p.sum()), float(p.abs().sum()), float(p[:: 997].double().sum())]
    # replicas must be identical
    t = torch.tensor(chk, dtype=torch.float64, device="cuda")
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi), ("replicas diverged", lo.tolist(), hi.tolist())
    return {"losses": losses, "checksum": chk, "transport": eng.stepper.sync.transport, "executor": type(eng.stepper).__name__}


def _replicas_equal(p):
    chk = torch.tensor([float(p.float().sum()), float(p.float().abs().sum()), float(p[::997].double().sum())], dtype=torch.float64, device="cuda")
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return bool(torch.equal(lo, hi)), chk.tolist()


def check_long_training(rank, world, transport, steps=60, use_peft=True):
    """`steps` updates with DP = world (through one ReLoRA merge + optimizer reset when use_peft): every loss, gradient norm and
    parameter must stay finite and the replicas bit-identical.  (Round-1 driver run: NaN after ~10 updates at 4 and 8 GPUs.)"""
    from relora_b200.engine.api import TrainingEngine
    from relora_b200.parallel.dist import init_distributed

    info = init_distributed("cuda", "nccl")
    kw = dict(use_peft=True, lora_r=128, relora=25, cycle_length=25, init_lora_a="kaiming") if use_peft else dict(use_peft=False)
    eng = TrainingEngine.build(info, model_config=os.path.join(ROOT, "configs", "llama_35m.json"), batch_size=4,
                               gradient_accumulation=2, total_batch_size=8 * world, max_length=256, scheduler="cosine_restarts" if use_peft else "cosine",
                               warmup_steps=5, restart_warmup_steps=2, lr=1e-3, num_training_steps=100, dtype="bfloat16", device="cuda",
                               comm=transport, seed=0, **kw)
    g = torch.Generator().manual_seed(1234 + rank)  # bench.py's per-rank token streams
    losses, norms = [], []
    for s in range(steps):
        ids = torch.randint(0, 32099, (2, 4, 256), generator=g).cuda()
        losses.append(float(eng.train_step_device(ids)))
        norms.append(float(eng.last_grad_norm))
        if (s + 1) % 10 == 0:
            p = eng.stepper.store.params
            assert bool(torch.isfinite(p.float()).all()), f"non-finite parameters after update {s}"
            same, chk = _replicas_equal(p)
            assert same, ("replicas diverged", s, chk)
    assert all(l == l and abs(l) < 1e4 for l in losses), losses
    assert all(n == n and n < 1e6 for n in norms), norms
    return {"losses": losses[::10] + [losses[-1]], "norms": norms[::10], "transport": eng.stepper.sync.transport,
            "executor": type(eng.stepper).__name__, "restarts": eng.n_lora_restarts, "steps": steps}


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mode = sys.argv[1]
    res = {}
    if mode == "allreduce":
        from relora_b200.parallel.symm import SymmComm

        comm = SymmComm()
        res = check_allreduce(comm, rank, world)
    elif mode.startswith("long_"):       # long_p2p | long_nccl
        res = check_long_training(rank, world, mode.split("_", 1)[1])
    elif mode.startswith("module_"):     # module_p2p | module_nccl: full-rank training on the module path
        res = check_long_training(rank, world, mode.split("_", 1)[1], steps=30, use_peft=False)
    elif mode.startswith("train_"):
        res = check_training(rank, world, mode.split("_", 1)[1])
    if rank == 0:
        print("RESULT " + json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
