"""Gradient all-reduce bus bandwidth: peer-memory kernels (P2P / NVLS multicast) vs NCCL, 64 KB - 64 MB (bf16).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 bench/allreduce_bench.py [--out f.json]

bus GB/s = 2 (N-1)/N * bytes / time (the usual all-reduce convention); device-timed, max over ranks.
"""
import argparse, json, os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from relora_b200.parallel.symm import SymmComm

ap = argparse.ArgumentParser(); ap.add_argument("--out", default=None); ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
comm = SymmComm()
rows = []

def timed(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / a.iters * 1e-3], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)

for kb in (64, 256, 1024, 4096, 16384, 65536):
    n = kb * 1024 // 2
    n = n // (8 * world) * 8 * world
    buf = comm.alloc(n, torch.bfloat16); buf.tensor.normal_()
    plain = torch.randn(n, device="cuda").to(torch.bfloat16)
    rec = {"kbytes": kb, "n_gpus": world}
    t = timed(lambda: dist.all_reduce(plain)); rec["nccl_us"] = t * 1e6
    comm.use_multicast = False
    t = timed(lambda: comm.all_reduce_(buf)); rec["p2p_us"] = t * 1e6
    if buf.mc_base:
        comm.use_multicast = True
        t = timed(lambda: comm.all_reduce_(buf)); rec["nvls_us"] = t * 1e6
    fac = 2 * (world - 1) / world * n * 2 / 1e9
    for k in ("nccl", "p2p", "nvls"):
        if f"{k}_us" in rec: rec[f"{k}_busGBs"] = fac / (rec[f"{k}_us"] * 1e-6)
    rows.append(rec)
    if rank == 0: print(json.dumps(rec), flush=True)
# ---- the fused data-parallel update of the training step (cast -> reduce-scatter + sum(g^2) -> norm/loss exchange -> AdamW on the
# owned shard -> parameter broadcast), llama_250m / llama_1b trainable sizes.  Roofline: each GPU receives (N-1)/N * 2 B per
# parameter in the reduce-scatter and sends the same in the broadcast; per-direction link bandwidth 770 GB/s measured (900 nominal).
for label, n_params in (("llama_250m", 98_888_448), ("llama_1b", 251_120_000)):
    n = n_params // (128 * world) * 128 * world
    for use_mc in (False, True):
        comm.use_multicast = use_mc
        pbuf = comm.alloc(n, torch.bfloat16); pbuf.tensor.normal_(std=0.02)
        gbuf = comm.alloc(n, torch.bfloat16)
        if use_mc and not pbuf.mc_base:
            continue
        grads = torch.randn(n, device="cuda") * 1e-3
        gred = torch.empty(n // world, device="cuda")
        m = torch.zeros(n // world, device="cuda", dtype=torch.bfloat16); v = torch.zeros_like(m)
        step = [0]
        def upd():
            step[0] += 1
            comm.fused_update(grads_f32=grads, grad_buf=gbuf, gred=gred, param_buf=pbuf, exp_avg=m, exp_avg_sq=v, n=n, lr=1e-4,
                              betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, step=step[0], max_norm=1.0, skip=None)
        t = timed(upd)
        link_bytes = (world - 1) / world * n * 2
        rec = {"fused_update": label, "n_gpus": world, "params": n, "transport": "nvls" if use_mc else "p2p", "us": t * 1e6,
               "link_GBs_per_direction": 2 * link_bytes / t / 1e9, "roofline_us_at_770GBs": 2 * link_bytes / 770e9 * 1e6,
               "fraction_of_roofline": (2 * link_bytes / 770e9) / t}
        rows.append(rec)
        if rank == 0: print(json.dumps(rec), flush=True)
        del pbuf, gbuf, grads, gred, m, v
comm.use_multicast = True
if rank == 0 and a.out:
    os.makedirs(os.path.dirname(a.out), exist_ok=True); json.dump(rows, open(a.out, "w"), indent=1)
dist.barrier(); dist.destroy_process_group()
