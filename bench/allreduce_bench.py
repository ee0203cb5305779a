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
if rank == 0 and a.out:
    os.makedirs(os.path.dirname(a.out), exist_ok=True); json.dump(rows, open(a.out, "w"), indent=1)
dist.barrier(); dist.destroy_process_group()
